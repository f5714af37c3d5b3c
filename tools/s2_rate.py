"""scqed_eigh_batched (HBM dense solver for n > 212) matrices/s for random real symmetric / complex Hermitian matrices
next to numpy.   python tools/s2_rate.py [n:batch ...]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from pycqed_b200 import engine
rng = np.random.default_rng(0)
cases = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(256, 2048), (300, 2048), (400, 1024), (600, 512), (1000, 128), (2000, 16), (3375, 3)]
for n, B in cases:
    a = rng.standard_normal((min(B, 4), n, n)); H = ((a + a.transpose(0, 2, 1)) / 2).astype(np.complex128)
    H = np.tile(H, (-(-B // len(H)), 1, 1))[:B]
    Hd = torch.as_tensor(H, device="cuda")
    for rep in range(2):
        Hc = Hd.clone(); torch.cuda.synchronize()
        engine.timing_enable(True); engine.timing_read(8)
        t = time.perf_counter()
        E, V, info = engine.eigh_batched(Hc, 10, want_vectors=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t
        red_ms, _ = engine.timing_read(8); engine.timing_enable(False)
    t = time.perf_counter(); w = np.linalg.eigvalsh(H[0].real); cpu = time.perf_counter() - t
    tf = (4.0 / 3.0) * n ** 3 * B / (red_ms * 1e-3) / 1e12 if red_ms > 0 else float("nan")
    print("n=%d batch=%d  GPU %.1f matrices/s  (reduction %.1f ms = %.2f TFLOP/s of real (4/3)n^3; numpy eigvalsh %.2f /s)  err %.1e  bad=%d" % (
        n, B, B / dt, red_ms, tf, 1 / cpu, np.abs(E.cpu().numpy()[0] - w[:10]).max(), int(info.sum())), flush=True)
