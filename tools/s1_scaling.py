#!/usr/bin/env python
"""S1 tridiagonalisation: device time per (matrix, column) against n, i.e. against the number of CTAs that fit one SM.
Tests whether the kernel is bound by the per-column dependent chain (time per column and CTA constant, throughput
proportional to resident CTAs) or by throughput.   python tools/s1_scaling.py [B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycqed_b200 import engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(0)
print("| n | smem/CTA KB (full storage) | CTAs/SM | tridiag ms | SM-clk per (matrix, column) | clk per column and CTA |")
print("|---|---|---|---|---|---|")
for n in (32, 48, 64, 72, 88, 101, 118, 128):
    A = rng.standard_normal((64, n, n)); A = A + A.transpose(0, 2, 1)
    H = torch.as_tensor(np.tile(A, (B // 64, 1, 1)).astype(np.complex128), device="cuda")
    for rep in range(2):
        Hc = H.clone()
        engine.timing_enable(True); engine.timing_read(1)
        E, V, info = engine.eigh_batched(Hc, 10, want_vectors=True, solver=1)
        torch.cuda.synchronize()
        ms, cnt = engine.timing_read(1); engine.timing_enable(False)
    smem = (n * n * 8 + 14 * n * 8 + 3000) / 1024.0
    ctas = min(int(227 // (smem + 1)), 8)
    clk = ms * 1e-3 * 1.965e9 * 148 / (H.shape[0] * (n - 1))
    print("| %d | %.0f | %d | %.3f | %.0f | %.0f |" % (n, smem, ctas, ms, clk, clk * ctas))
