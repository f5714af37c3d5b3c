#!/usr/bin/env python
"""Quick A/B of the S1 tridiagonalisation kernels on the bench workload: device time of the tagged kernel
(library CUDA events) and of the whole device-resident sweep, plus agreement of the eigenvalues.
  python tools/s1_bench.py [workload] [k]"""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycqed_b200 import engine
from pycqed_b200.plan import SweepPlan

name = sys.argv[1] if len(sys.argv) > 1 else "c2_fluxonium_res_10k"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
plan = SweepPlan.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", name + ".plan.json"))
grid = torch.as_tensor(plan.sweep_grid(), device="cuda")
res = "resonator" in plan.post
out = {}
for label, env in (("blk", {"SCQED_S1_BLK": "1"}), ("full", {"SCQED_S1_BLK": "0"})):
    os.environ.update(env)
    with engine.Engine(plan) as eng:
        for _ in range(3):
            E, V, _, _, info = eng.sweep_device(grid, k, want_vectors=True, resonator=res)
        torch.cuda.synchronize()
        engine.timing_enable(True)
        engine.timing_read(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            E, V, _, _, info = eng.sweep_device(grid, k, want_vectors=True, resonator=res)
        e1.record()
        torch.cuda.synchronize()
        ms, cnt = engine.timing_read(1)
        engine.timing_enable(False)
        out[label] = E.cpu().numpy()
        print("%-5s tridiag kernel %.3f ms/step   whole sweep %.3f ms/step   bad=%d" % (label, ms / 5, e0.elapsed_time(e1) / 5, int(info.sum())))
print("max |dE| blk vs full: %.3e (scale %.3g)" % (np.abs(out["blk"] - out["full"]).max(), np.abs(out["full"]).max()))
