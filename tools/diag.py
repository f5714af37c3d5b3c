"""Ad-hoc GPU diagnostics (not part of the test suite)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycqed_b200 import engine
from pycqed_b200.plan import SweepPlan
from oracle.plan_oracle import cluster_indices
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def load(name):
    g = dict(np.load(os.path.join(G, name + ".npz")))
    return g, SweepPlan.load(os.path.join(G, name + ".plan.json"))

def vec_diag(name):
    g, plan = load(name)
    k = int(g["k"]); pts = g["vec_points"]
    with engine.Engine(plan) as eng:
        res = eng.sweep(g["params"][pts], k, want_vectors=True)
        H = eng.assemble_dense(g["params"][pts]).cpu().numpy()
    for j in range(len(pts)):
        X = res.V[j].T; Vr = g["V"][j]; E = g["E"][pts[j]]
        print(name, "pt", pts[j], "params", g["params"][pts[j]], "info", res.info[j])
        print("  E    ", np.array2string(res.E[j], precision=10))
        print("  dE   ", np.array2string(res.E[j] - E, precision=2))
        R = H[j] @ X - X * res.E[j][None, :]
        print("  resid", np.array2string(np.abs(R).max(axis=0), precision=2))
        Rr = H[j] @ Vr - Vr * E[None, :]
        print("  refres", np.array2string(np.abs(Rr).max(axis=0), precision=2))
        for grp in cluster_indices(E, 1e-7, 1e-9):
            S = np.linalg.svd(X[:, grp].conj().T @ Vr[:, grp], compute_uv=False)
            print("  cluster", grp, "min sv", S.min())

def timeit(fn, reps=3):
    torch.cuda.synchronize(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

def perf():
    print("fp64 probe (DMMA, DFMA) TFLOP/s:", engine.probe_fp64(0))
    for name, k, vec, res, npts in [("c2_fluxonium_10k", 10, False, False, None), ("c2_fluxonium_10k", 10, True, False, None),
                         ("c2_fluxonium_res_10k", 20, True, True, None), ("c1_transmon", 10, False, False, None),
                         ("c4_averin_full", 10, False, False, None), ("c2_fluxonium_t200", 10, True, False, 2000),
                         ("c3_csfq4jj_t10", 10, True, False, 512), ("c5_fluxatom_t31_64", 10, True, False, 64),
                         ("c3_csfqfloat_t5", 10, True, False, 64), ("c3_csfq4jj_t27", 10, True, False, 8),
                         ("c3_csfqfloat_t7_full", 10, True, False, 8)]:
        plan = SweepPlan.load(os.path.join(G, name + ".plan.json"))
        grid = plan.sweep_grid()
        if npts is not None:
            reps = int(np.ceil(npts / len(grid)))
            grid = np.tile(grid, (reps, 1))[:npts]
        with engine.Engine(plan) as eng:
            p = torch.as_tensor(grid, device="cuda")
            ms = timeit(lambda: eng.sweep_device(p, k, want_vectors=vec, resonator=res), reps=2)
            t0 = time.time(); r = eng.sweep(grid, k, want_vectors=vec, resonator=res); t1 = time.time()
            print("%-24s n=%-6d pts=%-6d k=%d vec=%d res=%d  device %.2f ms (%.1f pts/s)  host-e2e %.1f ms (%.1f pts/s) bad=%d" % (
                name, plan.hilbert_dim, len(grid), k, vec, res, ms, len(grid) / ms * 1e3, (t1 - t0) * 1e3, len(grid) / (t1 - t0), int(r.info.sum())), flush=True)

if __name__ == "__main__":
    if "vec" in sys.argv: vec_diag("c3_csfqfloat_t5")
    if "perf" in sys.argv: perf()

def one(name, k, vec, npts, solver=0):
    plan = SweepPlan.load(os.path.join(G, name + ".plan.json"))
    grid = plan.sweep_grid()
    reps = int(np.ceil(npts / len(grid)))
    grid = np.tile(grid, (reps, 1))[:npts]
    with engine.Engine(plan) as eng:
        p = torch.as_tensor(grid, device="cuda")
        ms = timeit(lambda: eng.sweep_device(p, k, want_vectors=vec, solver=solver), reps=1)
        print("%s n=%d pts=%d device %.2f ms (%.2f pts/s)" % (name, plan.hilbert_dim, npts, ms, npts / ms * 1e3))

if __name__ == "__main__" and "one" in sys.argv:
    i = sys.argv.index("one")
    one(sys.argv[i + 1], int(sys.argv[i + 2]), bool(int(sys.argv[i + 3])), int(sys.argv[i + 4]))

def mf():
    for name, npts in [("c3_csfq4jj_t10", 1024), ("c3_csfqfloat_t5", 1024), ("c3_csfqfloat_t7_full", 1024), ("c3_csfq4jj_t27", 1024),
                       ("c3_csfqfloat_t10_full", 256), ("c3_csfq4jj_t50", 256), ("c5_fluxatom_t31_64", 64), ("c5_fluxatom_t180", 8)]:
        plan = SweepPlan.load(os.path.join(G, name + ".plan.json"))
        grid = plan.sweep_grid()
        reps = int(np.ceil(npts / len(grid)))
        grid = np.tile(grid, (reps, 1))[:npts]
        with engine.Engine(plan) as eng:
            p = torch.as_tensor(grid, device="cuda")
            l0 = engine.launch_count()
            ms = timeit(lambda: eng.sweep_device(p, 10, want_vectors=True, solver=4), reps=1)
            l1 = engine.launch_count()
            E, V, sc, post, info = eng.sweep_device(p, 10, want_vectors=True, solver=4)
            print("%-24s n=%-6d pts=%-5d matfree %.1f ms (%.1f pts/s) launches/run %d bad=%d" % (
                name, plan.hilbert_dim, npts, ms, npts / ms * 1e3, (l1 - l0) // 2, int((info != 0).sum().item())), flush=True)

if __name__ == "__main__" and "mf" in sys.argv:
    mf()

if __name__ == "__main__" and "mfone" in sys.argv:
    i = sys.argv.index("mfone")
    name, npts = sys.argv[i + 1], int(sys.argv[i + 2])
    plan = SweepPlan.load(os.path.join(G, name + ".plan.json"))
    grid = plan.sweep_grid()
    reps = int(np.ceil(npts / len(grid)))
    grid = np.tile(grid, (reps, 1))[:npts]
    with engine.Engine(plan) as eng:
        p = torch.as_tensor(grid, device="cuda")
        ms = timeit(lambda: eng.sweep_device(p, 10, want_vectors=True, solver=4), reps=1)
        print("%s n=%d pts=%d matfree %.1f ms" % (name, plan.hilbert_dim, npts, ms))
