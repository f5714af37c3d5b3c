import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from pycqed_b200 import engine
rng = np.random.default_rng(0)
for n, B in [(200, 2048), (256, 2048), (300, 2048), (400, 1024), (600, 512), (1000, 128)]:
    a = rng.standard_normal((B, n, n)); H = ((a + a.transpose(0, 2, 1)) / 2).astype(np.complex128)
    Hd = torch.as_tensor(H, device="cuda")
    for rep in range(2):
        Hc = Hd.clone(); torch.cuda.synchronize(); t = time.perf_counter()
        E, V, info = engine.eigh_batched(Hc, 10, want_vectors=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t
    t = time.perf_counter(); w = np.linalg.eigvalsh(H[:8].real); cpu = (time.perf_counter() - t) / 8
    print("n=%d batch=%d  GPU %.0f matrices/s   (numpy eigvalsh %.1f /s)  err %.1e" % (n, B, B / dt, 1 / cpu, np.abs(E.cpu().numpy()[:8] - w[:, :10]).max()))
