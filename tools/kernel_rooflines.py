"""Per-kernel achieved rates for the non-bench solvers (K2 assembly, S2 dense, S3 matrix-free), CUDA-event timed.
K2 is timed directly; for S2 / S3 run this script's `solve` mode under `ncu --metrics gpu__time_duration.sum`
and feed the launch list to tools/agg_launches.py.
usage: python tools/kernel_rooflines.py k2 | solve <plan> <points> <solver>"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycqed_b200 import engine
from pycqed_b200.plan import SweepPlan
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

if sys.argv[1] == "k2":
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6448.7) if os.path.exists("MEASURED_PEAKS.json") else 6448.7
    print("| plan | n | points | ms | GB/s written (16 n^2 B/pt) | frac of %.0f GB/s copy peak (read+write) |" % peak)
    print("|---|---|---|---|---|---|")
    for name, pts in [("c2_fluxonium_t200", 8192), ("c3_csfq4jj_t10", 4096), ("c5_fluxatom_t31_64", 2048), ("c3_csfqfloat_t7_full", 24), ("c3_csfq4jj_t50", 4)]:
        plan = SweepPlan.load(os.path.join(G, name + ".plan.json"))
        grid = plan.sweep_grid()
        grid = grid[np.arange(pts) % len(grid)]
        n = plan.hilbert_dim
        with engine.Engine(plan) as eng:
            p = torch.as_tensor(grid, device="cuda")
            H = torch.empty((pts, n, n), dtype=torch.complex128, device="cuda")
            lib, h = eng.lib, eng._handle
            def run():
                engine._check(lib.scqed_assemble_dense(h, p.data_ptr(), pts, 1, H.data_ptr(), engine._stream_ptr()))
            ms = timed(run)
        gb = pts * 16.0 * n * n / 1e9
        print("| %s | %d | %d | %.3f | %.0f | %.2f |" % (name, n, pts, ms, gb / (ms * 1e-3), gb / (ms * 1e-3) / peak))
else:
    name, pts, solver = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    plan = SweepPlan.load(os.path.join(G, name + ".plan.json"))
    grid = plan.sweep_grid()
    grid = grid[np.unique(np.linspace(0, len(grid) - 1, pts).astype(int))]
    with engine.Engine(plan) as eng:
        p = torch.as_tensor(grid, device="cuda")
        out = eng.sweep_device(p, 10, want_vectors=True, solver=solver); torch.cuda.synchronize()
        ms = timed(lambda: eng.sweep_device(p, 10, want_vectors=True, solver=solver), reps=1)
    print(name, "n", plan.hilbert_dim, "pts", len(grid), "solver", solver, "ms", ms, "bad", int((out[4] != 0).sum().item()))
