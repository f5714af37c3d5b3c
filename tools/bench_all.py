"""All BASELINE.json configs: GPU sweep throughput (device-resident and end-to-end through the host-buffer
C-ABI call) next to the CPU oracle port on a bounded sample.  Prints a markdown table."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycqed_b200 import engine
from pycqed_b200.plan import SweepPlan
from oracle.plan_oracle import PlanOracle, eval_program
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [  # name, plan, points (None = plan grid), k, vectors, resonator, cpu mode, cpu sample
    ("C1 transmon 40x25 phiZ x Qg", "c1_transmon", None, 10, True, False, "dense", 300),
    ("C4 Averin 101x201 Qc x phie (+2 scalar evaluables)", "c4_averin_full", None, 10, True, False, "dense", 400),
    ("C2 fluxonium t=101, 10k pts, k=10 values", "c2_fluxonium_10k", None, 10, False, False, "dense", 400),
    ("C2 fluxonium t=101, 10k pts, k=10 +vec +Resonator (bench)", "c2_fluxonium_res_10k", None, 10, True, True, "dense", 300),
    ("C2 fluxonium t=101, 10k pts, k=20 +vec +Resonator (example)", "c2_fluxonium_res_10k", None, 20, True, True, "dense", 200),
    ("C2 fluxonium t=200, 10k pts", "c2_fluxonium_t200", None, 10, True, False, "dense", 200),
    ("C3 CSFQ 3JJ 2-node t=10 (n=441), 1024 pts", "c3_csfq4jj_t10", 1024, 10, True, False, "dense", 60),
    ("C3 CSFQ 4JJ 2-node t=27 (n=3025), 16x64", "c3_csfq4jj_t27", None, 10, True, False, "dense", 3),
    ("C3 floating CSFQ 3-node t=7 (n=3375), 16x64", "c3_csfqfloat_t7_full", None, 10, True, False, "dense", 3),
    ("C3 floating CSFQ 3-node t=10 (n=9261), 256 pts", "c3_csfqfloat_t10_full", 256, 10, True, False, "sparse", 6),
    ("C3 CSFQ 4JJ 2-node t=50 (n=10201), 256 pts", "c3_csfq4jj_t50", 256, 10, True, False, "sparse", 4),
    ("C5 fluxonium-atom 31x31 (n=961), 64 pts", "c5_fluxatom_t31_64", None, 10, True, False, "dense", 20),
    ("C5 fluxonium-atom 180x180 (n=32400), 8 of 64 pts", "c5_fluxatom_t180", 8, 10, True, False, "none", 0),
]

def main():
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    only = sys.argv[1:] or None
    print("| config | n | points | GPU device pts/s | GPU e2e pts/s | CPU port pts/s (sample) | e2e / CPU |")
    print("|---|---|---|---|---|---|---|")
    for label, name, npts, k, vec, res, cpu, nsamp in CASES:
        if only and not any(o in name for o in only):
            continue
        plan = SweepPlan.load(os.path.join(G, name + ".plan.json"))
        grid = plan.sweep_grid()
        if npts is not None:
            idx = np.unique(np.linspace(0, len(grid) - 1, npts).astype(int)) if npts < len(grid) else np.arange(len(grid))
            grid = grid[idx]
        n = plan.hilbert_dim
        with engine.Engine(plan) as eng:
            p = torch.as_tensor(grid, device="cuda")
            eng.sweep_device(p, k, want_vectors=vec, resonator=res); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = eng.sweep_device(p, k, want_vectors=vec, resonator=res); e1.record(); torch.cuda.synchronize()
            dev = len(grid) / (e0.elapsed_time(e1) * 1e-3)
            bad = int((out[4] != 0).sum().item())
            hout = {"E": torch.empty((len(grid), k), dtype=torch.float64).pin_memory().numpy()}
            if vec:
                hout["V"] = torch.empty((len(grid), k, n), dtype=torch.complex128).pin_memory().numpy()
            eng.sweep(grid, k, want_vectors=vec, resonator=res, out=hout)
            t0 = time.perf_counter(); r = eng.sweep(grid, k, want_vectors=vec, resonator=res, out=hout); e2e = len(grid) / (time.perf_counter() - t0)
        cpu_rate, samp = None, ""
        if cpu != "none":
            orc = PlanOracle(plan)
            sidx = np.unique(np.linspace(0, len(grid) - 1, nsamp).astype(int))
            t0 = time.perf_counter()
            if cpu == "dense":
                ref = orc.sweep(grid[sidx], k, get_vectors=vec, resonator=res)
            else:   # the reference example's shift-invert setting (examples/csfq-example.py:513-519)
                ref = orc.sweep(grid[sidx], k, get_vectors=vec, sparse_opts={"sigma": -300, "mode": "normal", "maxiter": None, "tol": 0})
            dt = time.perf_counter() - t0
            cpu_rate = len(sidx) / dt
            err = np.max(np.abs(r.E[sidx] - ref["E"]) / np.maximum(np.abs(ref["E"]), 1e-3 * np.abs(ref["E"]).max()))
            samp = "%d pts, %.1f s, %s, max rel dE %.1e" % (len(sidx), dt, cpu, err)
        print("| %s | %d | %d | %.4g | %.4g | %s | %s |%s" % (
            label, n, len(grid), dev, e2e, ("%.4g (%s)" % (cpu_rate, samp)) if cpu_rate else "-", ("%.0fx" % (e2e / cpu_rate)) if cpu_rate else "-",
            " nonconverged=%d" % bad if bad else ""), flush=True)

if __name__ == "__main__":
    main()
