import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycqed_b200 import engine
from pycqed_b200.plan import SweepPlan
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
plan = SweepPlan.load(os.path.join(G, "c2_fluxonium_res_10k.plan.json"))
grid = plan.sweep_grid(); n = plan.hilbert_dim; k = 10; npts = len(grid)
hE = torch.empty((npts, k), dtype=torch.float64).pin_memory()
hV = torch.empty((npts, k, n), dtype=torch.complex128).pin_memory()
hR = torch.empty((npts, 111, k), dtype=torch.float64).pin_memory()
out = {"E": hE.numpy(), "V": hV.numpy(), "Erwa": hR.numpy()}
with engine.Engine(plan) as eng:
    p = torch.as_tensor(grid, device="cuda")
    for _ in range(3): eng.sweep_device(p, k, want_vectors=True, resonator=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): eng.sweep_device(p, k, want_vectors=True, resonator=True)
    torch.cuda.synchronize(); print("device  %.2f ms" % ((time.perf_counter() - t0) * 100))
    for ch in [1, 2, 3, 4, 6, 8]:
        os.environ["SCQED_HOST_CHUNKS"] = str(ch)
        for _ in range(2): eng.sweep(grid, k, want_vectors=True, resonator=True, out=out)
        t0 = time.perf_counter()
        for _ in range(10): eng.sweep(grid, k, want_vectors=True, resonator=True, out=out)
        print("chunks %d: e2e %.2f ms" % (ch, (time.perf_counter() - t0) * 100))
    # raw D2H bandwidth
    d = torch.empty(hV.shape, dtype=torch.complex128, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): hV.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("D2H pinned %.1f GB/s" % (hV.numel() * 16 / dt / 1e9))
