#!/usr/bin/env python
"""bench.py -- sweep-points/s of the hot path (assemble H(p) + lowest-k eigenpairs + resonator
post-op) on N B200s, with the reference's CPU path timed beside it.

Workload (BASELINE.json configs[1]): fluxonium qubit A coupled to a readout resonator
(examples/fluxonium-example.py:25-73), oscillator basis truncation 101 (n = 101), 10 000-point
external-flux sweep phiZ in [-1, 1], lowest 10 eigenpairs (values + vectors) and the `Resonator`
evaluable (RWA strips -> dispersive shift), synthetic deterministic linspace grid.
A "step" is one pass of that sweep.  Other workloads: --workload <plan name in tests/golden>.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); sweep points are sharded, every rank runs the
same number of points (weak scaling: `value`, `e2e`), the only collective is the final NCCL gather of the
eigenvalues; `strong_scaling` is the same 10 000-point sweep cut over the N ranks.  At N = 1 the line also carries
a `workloads` block: the CSFQ / Averin / fluxonium-atom configs of BASELINE.json, each with device-resident and
end-to-end points/s, the roofline of its dominant kernel, and the CPU arm on a sample of the same points.  `--impl reference` times the CPU oracle port of the reference's per-point loop
(the reference is pure Python and is not present on the GPU box; see oracle/plan_oracle.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

METRIC = "diagonalised sweep points/sec (lowest-10 eigenpairs)"
# `ncu --set full` capture the headline kernel's DRAM traffic is read from, and the points of that launch
TRAFFIC_CSV, TRAFFIC_PTS = "r02_s1t_tiled_ncu_summary.csv", 10000.0
UNIT = "points/s"


def load_workload(name):
    from pycqed_b200.plan import SweepPlan
    return SweepPlan.load(os.path.join(GOLDEN, name + ".plan.json"))


def workload_config(plan, name, k, vectors, resonator, n_gpus):
    return {
        "workload": "%s: fluxonium+resonator, oscillator basis n=%d, %d-point phiZ sweep per GPU, k=%d, "
                    "vectors=%s, Resonator(nmax=100)=%s" % (name, plan.hilbert_dim, len(plan.sweep_grid()), k,
                                                           bool(vectors), bool(resonator)),
        "n": plan.hilbert_dim, "points_per_gpu": int(len(plan.sweep_grid())), "k": k,
        "vectors": bool(vectors), "resonator": bool(resonator),
        "sharding": "contiguous blocks of the flattened grid, %d rank(s), no data-path collective; "
                    "NCCL all_gather of eigenvalues at the end of each step" % n_gpus,
        "l2": "explicit 256 MiB L2 flush between timed steps (outside the per-step event pair); per-step outputs "
              "(eigenvectors + strips, ~250 MB) also exceed the 126 MB L2",
    }


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's per-point loop)
# ------------------------------------------------------------------------------------------------
def cpu_baseline(plan, k, vectors, resonator, budget_s=12.0, max_pts=2000):
    """Times oracle.plan_oracle.PlanOracle.sweep -- coefficient evaluation, CSR Kronecker assembly
    (numerical_system.py:509-594), scipy.linalg.eigh(subset_by_index) (util.py:153) and the RWA
    strips (numerical_system.py:736-784) -- point by point like the reference's loop
    (numerical_system.py:944-965), on a bounded sample of the same grid."""
    from oracle.plan_oracle import PlanOracle
    try:        # torchrun exports OMP_NUM_THREADS=1; the baseline may use every host core
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    grid = plan.sweep_grid()
    orc = PlanOracle(plan)
    # calibrate on a few points, then size the sample for ~budget_s
    orc.sweep(grid[:2], k, get_vectors=vectors, resonator=resonator)      # warm caches / imports
    t0 = time.perf_counter()
    orc.sweep(grid[:6], k, get_vectors=vectors, resonator=resonator)
    per = (time.perf_counter() - t0) / 6
    n = int(max(8, min(max_pts, len(grid), budget_s / max(per, 1e-6))))
    idx = np.unique(np.linspace(0, len(grid) - 1, n).astype(int))
    t0 = time.perf_counter()
    orc.sweep(grid[idx], k, get_vectors=vectors, resonator=resonator)
    dt = time.perf_counter() - t0
    try:
        import threadpoolctl
        threads = max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [1])
    except Exception:
        threads = os.cpu_count() or 1
    return {
        "value": len(idx) / dt, "unit": UNIT, "cores": int(threads), "kind": "port",
        "host_cpus": os.cpu_count(),
        "sample": "%d of %d grid points (evenly spaced), %.1f s; per-point loop as numerical_system.py:944-965 "
                  "(single Python thread, LAPACK/BLAS threads = cores); coefficients by the compiled program, "
                  "i.e. WITHOUT the reference's 2-30 ms/pt sympy substitution, so this baseline is faster than "
                  "the reference itself" % (len(idx), len(grid), dt),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    plan = load_workload(args.workload)
    k, vectors, resonator = args.k, not args.no_vectors, (not args.no_resonator) and "resonator" in plan.post
    times, vals = [], []
    base = None
    # each step is a bounded sample of the workload; the whole run stays within ~2 minutes
    budget = min(args.ref_budget, 120.0 / max(1, args.warmup + args.steps))
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(plan, k, vectors, resonator, budget_s=budget)
        if i >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * len(plan.sweep_grid()) / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128 H)", "data": "synthetic",
        "config": workload_config(plan, args.workload, k, vectors, resonator, 1),
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def _par_worker(args):
    """One process of the parallel CPU baseline: its share of the sample with ONE BLAS thread."""
    workload, k, vectors, resonator, pts = args
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=1)
    except Exception:
        pass
    from oracle.plan_oracle import PlanOracle
    plan = load_workload(workload)
    orc = PlanOracle(plan)
    orc.sweep(pts[:1], k, get_vectors=vectors, resonator=resonator)
    t0 = time.perf_counter()
    orc.sweep(pts, k, get_vectors=vectors, resonator=resonator)
    return time.perf_counter() - t0


def cpu_baseline_parallel(workload, plan, k, vectors, resonator, per_point_s, budget_s=8.0):
    """A STRONGER baseline than the reference itself (which is one Python loop): the oracle port with one
    process per host core, one BLAS thread each ("parallel reference").  Reported beside
    `cpu_baseline`, never instead of it."""
    import multiprocessing as mp
    procs = max(1, os.cpu_count() or 1)
    grid = plan.sweep_grid()
    # a single-threaded worker is somewhat slower per point than the multi-threaded loop that was timed
    n_each = int(max(4, min(len(grid) // procs, budget_s / max(per_point_s * 1.5, 1e-6))))
    idx = np.unique(np.linspace(0, len(grid) - 1, n_each * procs).astype(int))
    shares = [grid[idx[i::procs]] for i in range(procs)]
    saved = {v: os.environ.get(v) for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    try:
        for v in saved:                       # spawned workers read these when they load numpy / scipy
            os.environ[v] = "1"
        ctx = mp.get_context("spawn")
        t0 = time.perf_counter()
        with ctx.Pool(procs) as pool:
            res = pool.map_async(_par_worker, [(workload, k, vectors, resonator, sh) for sh in shares])
            times = res.get(timeout=120)
        wall = time.perf_counter() - t0
        busy = max(times)
        return {"value": len(idx) / busy, "unit": UNIT, "processes": procs, "kind": "port, one process per core",
                "sample": "%d points over %d processes, slowest worker %.1f s (pool wall %.1f s incl. start-up)" % (len(idx), procs, busy, wall),
                "note": "not how the reference runs (it is a single Python loop); shown as the strongest CPU figure"}
    except Exception as exc:
        return {"error": str(exc)[:200]}
    finally:
        for v, old in saved.items():
            if old is None:
                os.environ.pop(v, None)
            else:
                os.environ[v] = old


# ------------------------------------------------------------------------------------------------
# The other BASELINE configs (CSFQ, Averin, fluxonium atom): one block per workload on the bench line
# ------------------------------------------------------------------------------------------------
# (plan in tests/golden, points (None = the plan's grid), CPU arm, CPU sample points, BASELINE config)
WORKLOADS = [
    ("c1_transmon", None, "dense", 250, "configs[0] transmon, charge basis n=81, 40x25 phiZ x Qg"),
    ("c4_averin_full", None, "dense", 400, "configs[3] Averin coupler, charge basis n=41, 101x201 Qc x phie, 2 scalar evaluables"),
    ("c3_csfqfloat_t7_full", None, "dense", 3, "configs[2] floating CSFQ 3-node charge t=7 (n=3375), 16x64 asymmetry x flux"),
    ("c3_csfqfloat_t10_full", 256, "shift-invert", 8, "configs[2] floating CSFQ 3-node charge t=10 (n=9261), 256 of 16x64"),
    ("c3_csfq4jj_t50", 256, "shift-invert", 6, "configs[2] CSFQ 4JJ 2-node charge t=50 (n=10201), 256 of 16x64"),
    ("c5_fluxatom_t31_64", None, "dense", 24, "configs[4] fluxonium atom 31x31 (n=961), 64-point phi- sweep"),
    ("c5_fluxatom_t180", 8, "arpack", 1, "configs[4] fluxonium atom 180x180 (n=32400), 8 of 64 points"),
]
_SHIFT_INVERT = {"sigma": -300, "mode": "normal", "maxiter": None, "tol": 0}     # examples/csfq-example.py:513-519
_ARPACK_EXACT = {"sigma": None, "mode": "normal", "maxiter": None, "tol": 0, "which": "SA"}   # numerical_system.py:790 with tol=0


def _subgrid(grid, npts):
    if npts is None or npts >= len(grid):
        return grid
    return grid[np.unique(np.linspace(0, len(grid) - 1, npts).astype(int))]


def workload_block(name, npts, cpu_mode, cpu_sample, label, k, dev, peak_fp64, hbm_peak, budget_s):
    """Device-resident and end-to-end throughput of one more workload, the per-kernel roofline of its dominant
    kernel (library CUDA-event spans + work counters), and the CPU arm (oracle port) on a bounded sample of the
    same points with the eigenvalue agreement of that sample."""
    import torch
    from pycqed_b200 import engine
    from oracle.plan_oracle import PlanOracle, eig_rel_err
    plan = load_workload(name)
    grid = _subgrid(plan.sweep_grid(), npts)
    n, n_pts = plan.hilbert_dim, len(grid)
    blk = {"workload": name, "config": label, "n": n, "points": n_pts, "k": k, "vectors": True}
    t0 = time.perf_counter()
    with engine.Engine(plan, device=dev.index) as eng:
        blk["plan_upload_s"] = time.perf_counter() - t0
        p = torch.as_tensor(grid, device=dev)
        eng.sweep_device(p, k, want_vectors=True)
        torch.cuda.synchronize(dev)
        engine.timing_enable(True)
        for tag in range(1, 9):
            engine.timing_read(tag)
        engine.counter_read(1), engine.counter_read(2)
        l0 = engine.launch_count()
        reps = 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            E, V, scal, post, info = eng.sweep_device(p, k, want_vectors=True)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        spans = {tag: engine.timing_read(tag) for tag in range(1, 9)}
        engine.timing_enable(False)
        matvecs, mf_flops = engine.counter_read(1) / reps, engine.counter_read(2) / reps
        blk["device"] = {"value": n_pts / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                         "gpu_launches": int((engine.launch_count() - l0) / reps)}
        blk["nonconverged_points"] = int((info != 0).sum().item())
        # end to end: host params in, E + eigenvectors out to page-locked host arrays, through scqed_sweep_run_host
        hout = {"E": engine.pinned_empty((n_pts, k), np.float64), "V": engine.pinned_empty((n_pts, k, n), np.complex128)}
        eng.sweep(grid, k, want_vectors=True, out=hout)
        t0 = time.perf_counter()
        r = eng.sweep(grid, k, want_vectors=True, out=hout)
        dt = time.perf_counter() - t0
        blk["e2e"] = {"value": n_pts / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": int(grid.nbytes),
                      "d2h_bytes_per_step": int(hout["E"].nbytes + hout["V"].nbytes + n_pts * (4 + 8 * max(1, plan.program.n_scalar)))}
        E_gpu = r.E.copy()
    # committed reference outputs of the same circuit at this truncation (tests/golden/<name>s.npz, minted by the
    # unmodified reference): solve the fixture's own points as well
    E_gold, ref_spp = None, None
    gold_name = os.path.join(GOLDEN, name + "s.npz")
    if os.path.exists(gold_name):
        gold = np.load(gold_name, allow_pickle=False)
        with engine.Engine(load_workload(name + "s"), device=dev.index) as geng:
            E_gold = (geng.sweep(np.asarray(gold["params"], dtype=np.float64), k).E, np.asarray(gold["E"])[:, :k])
        ref_spp = float(gold["ref_seconds_per_point"]) if "ref_seconds_per_point" in gold.files else None
    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    filt_ms = sum(spans[t][0] for t in (5, 6, 7)) / reps
    if filt_ms > 0:
        form = max((5, 6, 7), key=lambda t: spans[t][0])
        kern = {5: "mf_cheb_stencil (whole Chebyshev filter per launch, vector slices in shared memory / DSMEM)",
                6: "mf_cheb_smem / mf_cheb_smem_d2 (whole Chebyshev filter per launch, vectors in shared memory)",
                7: "mf_d2_dmma_kernel / mf_cheb_step (one mat-vec per launch, vectors streamed from L2/HBM)"}[form]
        tf = mf_flops / (filt_ms * 1e-3) / 1e12
        roof = {"kernel": kern, "bound": "tensor" if form == 7 and n > 20000 else "fp64", "kernel_ms_per_step": filt_ms,
                "kernel_share_of_step": filt_ms / ms, "vector_matvecs_per_step": matvecs,
                "us_per_vector_matvec": 1e3 * filt_ms / max(matvecs, 1), "flops_per_vector_matvec": mf_flops / max(matvecs, 1),
                "achieved": tf, "peak": peak_fp64, "unit": "TFLOP/s", "frac": tf / peak_fp64,
                "note": "algorithmic flops of the Kronecker mat-vec (SURVEY.md 8d: 8 n sum_t prod_k w_tk per vector; the combined "
                        "two-node form n (d0 + d1) real MACs) x mat-vecs issued / summed filter spans (CUDA events on the "
                        "launching stream); peak = FP64 rate probed in this run"}
        if form == 7:
            gbs = 32.0 * n * matvecs / (filt_ms * 1e-3) / 1e9
            roof["stream_gbs"] = gbs
            roof["stream_frac_of_hbm_peak"] = gbs / hbm_peak if hbm_peak else None
    elif spans[4][0] > 0:
        kms = spans[4][0] / reps
        by = n_pts * (k * n * 16.0 + k * 8.0 + 40.0 * n)
        roof = {"kernel": "s1_direct_tridiag_kernel (Hermitian tridiagonal plan: no Householder stage) + bisection / inverse iteration / phase back-transform",
                "bound": "hbm", "kernel_ms_per_step": kms, "kernel_share_of_step": kms / ms,
                "achieved": by / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": (by / (ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                "note": "whole step against the result bytes it must write (k eigenvectors of n complex128 + k eigenvalues + 40 n B of "
                        "tridiagonal workspace per point); latency bound at this size"}
    elif spans[1][0] + spans[2][0] > 0:
        real = spans[1][0] >= spans[2][0]
        kms = (spans[1][0] + spans[2][0]) / reps
        tf = (4.0 / 3.0 if real else 16.0 / 3.0) * n ** 3 * n_pts / (kms * 1e-3) / 1e12
        roof = {"kernel": "s1_tridiag_kernel", "bound": "fp64", "kernel_ms_per_step": kms, "kernel_share_of_step": kms / ms,
                "achieved": tf, "peak": peak_fp64, "unit": "TFLOP/s", "frac": tf / peak_fp64}
    elif spans[8][0] > 0:
        kms = spans[8][0] / reps
        tf = (16.0 / 3.0) * n ** 3 * n_pts / (kms * 1e-3) / 1e12
        roof = {"kernel": "dense (HBM) tridiagonal reduction", "bound": "tensor", "kernel_ms_per_step": kms, "kernel_share_of_step": kms / ms,
                "achieved": tf, "peak": peak_fp64, "unit": "TFLOP/s", "frac": tf / peak_fp64}
    else:
        roof = None
    blk["roofline"] = roof
    # ---- CPU arm on a bounded sample of the same points -----------------------------------------------------
    if E_gold is not None:
        blk["max_rel_dE_vs_reference_output"] = {
            "value": eig_rel_err(*E_gold), "points": len(E_gold[0]),
            "fixture": "tests/golden/%ss.npz: the unmodified reference's sparse path (util.py:76-125, tol=0)" % name,
            "reference_seconds_per_point_when_minted": ref_spp}
    orc = PlanOracle(plan)
    sidx = np.unique(np.linspace(0, n_pts - 1, cpu_sample).astype(int)) if cpu_sample > 1 else np.array([n_pts // 2])
    sparse = None if cpu_mode == "dense" else dict(_SHIFT_INVERT if cpu_mode == "shift-invert" else _ARPACK_EXACT)
    t0 = time.perf_counter()
    done, Eref = [], []
    for i in sidx:                                  # stop early when the sample outgrows its budget
        Eref.append(orc.sweep(grid[i:i + 1], k, get_vectors=True, sparse_opts=sparse)["E"][0])
        done.append(i)
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    blk["cpu"] = {"value": len(done) / dt, "unit": UNIT, "kind": "port", "cores": os.cpu_count(),
                  "sample": "%d of the %d points, %.1f s, %s" % (len(done), n_pts, dt,
                            "scipy.linalg.eigh(subset_by_index) as util.py:153" if cpu_mode == "dense" else
                            "eigsh(sigma=-300, tol=0) shift-invert, the reference example's setting (examples/csfq-example.py:513-519)"
                            if cpu_mode == "shift-invert" else
                            "eigsh(which='SA', tol=0) on the CSR matrix (util.py:76-125; dense is impossible: 16.8 GB per matrix)")}
    blk["max_rel_dE"] = eig_rel_err(E_gpu[np.asarray(done)], np.asarray(Eref))
    blk["dE_reference"] = "CPU arm on the sampled points"
    blk["e2e_over_cpu"] = blk["e2e"]["value"] / blk["cpu"]["value"]
    return blk


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

def dense_eigh_block(dev, hbm_peak):
    """`diagonalize()` of dense real symmetric matrices beyond the shared-memory path (reference: util.py:127-176 -> LAPACK,
    one matrix at a time): scqed_eigh_batched on batches of seeded random matrices, lowest 10 eigenpairs with vectors.
    matrices/s from CUDA events around the whole call (inputs resident in HBM), the reduction's share from the library's own
    event span (tag 8), numpy.linalg.eigh on the host cores beside it, eigenvalues checked against it."""
    import numpy as np, torch
    from pycqed_b200 import engine
    out = []
    rng = np.random.default_rng(7)
    for n, B in ((400, 1184), (1000, 296), (2000, 148)):
        a = rng.standard_normal((2, n, n))
        Hs = ((a + a.transpose(0, 2, 1)) / 2).astype(np.complex128)
        H = torch.as_tensor(np.tile(Hs, (B // 2, 1, 1)), device=dev)
        ms, red = [], []
        for rep in range(3):
            Hc = H.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            engine.timing_enable(True); engine.timing_read(8)
            e0.record()
            E, V, info = engine.eigh_batched(Hc, 10, want_vectors=True)
            e1.record(); torch.cuda.synchronize(dev)
            r, _ = engine.timing_read(8); engine.timing_enable(False)
            ms.append(e0.elapsed_time(e1)); red.append(r)
        t = min(ms[1:]); rms = min(red[1:])
        t0 = time.perf_counter(); w, v = np.linalg.eigh(Hs[0].real); cpu_s = time.perf_counter() - t0
        Eh = E.cpu().numpy()
        # one-stage tridiagonalisation on the packed lower triangle: stored half read once per column, update written once per
        # panel of NB columns => (4/3)(1 + 1/NB) n^3 B per matrix (the streamed kernel, n >= 560); below that the register
        # kernel reads every tile twice and is latency bound, no HBM roofline quoted
        nb = 4 if n < 1400 else 8                     # what run_dense_tiled_batch picks at these sizes (panels in shared memory / in L2)
        alg = (4.0 / 3.0) * (1.0 + 1.0 / nb) * n ** 3 * B
        blk = {"workload": "eigh_batched real symmetric n=%d, %d matrices, k=10 + vectors" % (n, B), "n": n, "batch": B,
               "value": B / (t * 1e-3), "unit": "matrices/s", "ms": t, "reduction_ms": rms,
               "cpu": {"value": 1.0 / cpu_s, "unit": "matrices/s", "kind": "numpy.linalg.eigh (LAPACK, all host threads), 1 matrix"},
               "max_rel_dE": float(np.abs(Eh[0] - w[:10]).max() / np.abs(w).max()), "nonconverged": int(info.sum().item())}
        if n >= 560 and rms > 0:
            blk["roofline"] = {"bound": "hbm", "kernel": "dtc_tridiag_kernel", "achieved": alg / (rms * 1e-3) / 1e9, "peak": hbm_peak,
                               "unit": "GB/s", "frac": alg / (rms * 1e-3) / 1e9 / hbm_peak,
                               "note": "algorithmic bytes (4/3)(1 + 1/NB) n^3 B per matrix, NB = %d / the reduction's event span" % nb}
        out.append(blk)
    # the same path behind the sweep API: fluxonium plan (c2_fluxonium_t200) at node truncation 640, dense real symmetric
    # H(phiZ) assembled by K2 into HBM and reduced by S2t; end to end through Engine.sweep (host arrays in and out)
    try:
        from pycqed_b200.plan import SweepPlan
        d = json.load(open(os.path.join(ROOT, "tests", "golden", "c2_fluxonium_t200.plan.json")))
        d["nodes"][0]["trunc"] = 640
        plan = SweepPlan.from_dict(d)
        full = plan.sweep_grid()
        grid = np.ascontiguousarray(full[np.linspace(0, len(full) - 1, 592).round().astype(int)])
        with engine.Engine(plan, device=dev.index) as eng:
            ts = []
            for rep in range(3):
                t0 = time.perf_counter(); res = eng.sweep(grid, 10, want_vectors=True); ts.append(time.perf_counter() - t0)
        t = min(ts[1:])
        from oracle.plan_oracle import PlanOracle
        t0 = time.perf_counter(); ref = PlanOracle(plan).sweep(grid[:4], 10, get_vectors=True); cpu_s = (time.perf_counter() - t0) / 4
        out.append({"workload": "c2 fluxonium at truncation 640 (dense real symmetric n=640), 592-point phiZ sweep, k=10 + vectors",
                    "n": 640, "points": 592, "e2e_pts_s": 592 / t, "ms": t * 1e3,
                    "cpu": {"value": 1.0 / cpu_s, "unit": "points/s", "kind": "oracle port, 4 points"},
                    "max_rel_dE": float(np.abs(res.E[:4] - ref["E"]).max() / np.abs(ref["E"]).max()),
                    "nonconverged": int(np.count_nonzero(res.info))})
    except Exception as exc:
        out.append({"workload": "c2 fluxonium at truncation 640", "error": "%s: %s" % (type(exc).__name__, str(exc)[:300])})
    return out

def run_ours(args):
    import torch
    import torch.distributed as dist
    from pycqed_b200 import engine
    from pycqed_b200.distributed import sweep_sharded, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    plan = load_workload(args.workload)
    k, vectors = args.k, not args.no_vectors
    resonator = (not args.no_resonator) and "resonator" in plan.post
    base_grid = plan.sweep_grid()
    n_pts, n = len(base_grid), plan.hilbert_dim
    # weak scaling: the global grid has N x n_pts points (an N-times finer sweep of the same axis), rank r owns the
    # contiguous block [r n_pts, (r+1) n_pts) -- exactly what distributed.sweep_sharded gives it
    if world > 1 and len(plan.sweep_specs) == 1:
        s = plan.sweep_specs[0]
        weak_grid = np.linspace(s["start"], s["end"], n_pts * world)[:, None].copy()
    else:
        weak_grid = np.concatenate([base_grid] * world, axis=0)
    lo, hi = shard_bounds(len(weak_grid), world, rank)
    grid = np.ascontiguousarray(weak_grid[lo:hi])

    eng = engine.Engine(plan, device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ns = plan.program.n_scalar
    nstrips = plan.post["resonator"]["nmax"] + 1 + k if resonator else 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def time_device(pts, steps, warmup):
        """K steps of the device-resident sweep of this rank's `pts` (+ the NCCL gather of the eigenvalues), each
        bracketed by barrier + synchronize, CUDA events on the launching stream; returns (sum of ms, max over ranks)."""
        p_dev = torch.as_tensor(pts, device=dev)
        gathered = [torch.empty((len(pts), k), dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None
        out = None

        def step():
            E, V, scal, post, info = eng.sweep_device(p_dev, k, want_vectors=vectors, resonator=resonator)
            if world > 1:
                dist.all_gather(gathered, E)
            return info
        for _ in range(warmup):
            step()
            flush.fill_(1)
        barrier()
        times = []
        for _ in range(steps):
            flush.fill_(0)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = step()
            e1.record()
            barrier()
            times.append(e0.elapsed_time(e1))
        return max_over_ranks(float(np.sum(times))), out

    def time_e2e(global_grid, steps, mode, real_out):
        """The product path for a multi-process run: distributed.sweep_sharded -> Engine.sweep ->
        scqed_sweep_run_host on this rank's block, host parameters in, results out to host arrays, NCCL gather of
        eigenvalues / scalars / info at the end.  mode: 'pinned' (page-locked result arrays), 'pageable' (plain
        numpy arrays: the library stages them through its own pinned ring), 'lazy' (eigenvectors stay on the
        device, fetched on demand -- what paramSweep does when only Resonator consumes them)."""
        a, b = shard_bounds(len(global_grid), world, rank)
        m = b - a
        alloc = engine.pinned_empty if mode != "pageable" else (lambda shape, dt: np.empty(shape, dtype=dt))
        out = {"E": alloc((m, k), np.float64)}
        if vectors and mode != "lazy":
            out["V"] = alloc((m, k, n), np.float64 if real_out else np.complex128)
        if resonator:
            out["Erwa"] = alloc((m, nstrips, k), np.float64)
        kw = dict(out=out, gather_strips=False)
        if mode == "lazy":
            kw["vectors_on_device"] = True
        elif vectors:
            kw["real_vectors"] = real_out

        def step():
            return sweep_sharded(eng, global_grid, k, want_vectors=vectors, resonator=resonator, device=dev, **kw)
        step()
        barrier()
        ts = []
        for _ in range(steps):
            flush.fill_(0)
            barrier()
            t0 = time.perf_counter()
            step()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        d2h = sum(a_.nbytes for a_ in out.values()) + m * max(ns, 1) * 8 + m * 4
        return max_over_ranks(float(np.mean(ts))), int(d2h)

    # ---- device-resident headline ---------------------------------------------------------------------
    l0 = engine.launch_count()
    engine.timing_enable(True)
    engine.timing_read(1), engine.timing_read(2)           # clear
    with ClockSampler(local) as clocks:
        total_ms, info = time_device(grid, args.steps, args.warmup)
    engine.timing_enable(False)
    k_real_ms, k_real_n = engine.timing_read(1)
    k_cplx_ms, k_cplx_n = engine.timing_read(2)
    warm_frac = args.steps / float(args.steps + args.warmup)     # spans and launches also cover the warm-up steps
    k_real_ms, k_cplx_ms = k_real_ms * warm_frac, k_cplx_ms * warm_frac
    launches = int(round((engine.launch_count() - l0) * warm_frac))
    bad = int(info.sum().item())
    ms_per_step = total_ms / args.steps
    value = world * n_pts / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (what paramSweep does) -------------------
    e2e_steps = max(1, min(args.steps, 10))
    real_out = False
    if vectors and not args.complex_vectors:
        probe = eng.sweep(grid[:64], k, want_vectors=True, resonator=resonator, real_vectors="auto")
        real_out = probe.V.dtype == np.float64
    e2e_s, d2h = time_e2e(weak_grid, e2e_steps, "pinned", real_out)
    e2e_page_s, _ = time_e2e(weak_grid, e2e_steps, "pageable", real_out)
    e2e_lazy_s, d2h_lazy = (time_e2e(weak_grid, e2e_steps, "lazy", real_out) if vectors else (None, None))
    h2d = int(grid.nbytes)

    # ---- D2H bandwidth probe: every rank copies its per-step result bytes to page-locked host memory at once -----
    hb = torch.empty(d2h, dtype=torch.uint8).pin_memory()
    db = torch.empty(d2h, dtype=torch.uint8, device=dev)
    hb.copy_(db)
    barrier()
    pt = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        hb.copy_(db, non_blocking=True)
        torch.cuda.synchronize()
        pt.append(time.perf_counter() - t0)
    probe_s = max_over_ranks(min(pt))
    del hb, db

    # ---- strong scaling: the SAME 10 000-point sweep cut over the N ranks ---------------------------------
    strong = None
    if world > 1:
        a, b = shard_bounds(n_pts, world, rank)
        st_steps = max(1, min(args.steps, 10))
        st_ms, _ = time_device(np.ascontiguousarray(base_grid[a:b]), st_steps, 2)
        st_e2e_s, _ = time_e2e(base_grid, st_steps, "pinned", real_out)
        st_lazy_s, _ = time_e2e(base_grid, st_steps, "lazy", real_out) if vectors else (None, None)
        strong = {"points_total": n_pts, "points_per_gpu": b - a, "value": n_pts / (st_ms / st_steps * 1e-3), "unit": UNIT,
                  "ms_per_step": st_ms / st_steps, "e2e_value": n_pts / st_e2e_s, "e2e_ms_per_step": st_e2e_s * 1e3,
                  "e2e_lazy_vectors_value": (n_pts / st_lazy_s) if st_lazy_s else None,
                  "note": "the BASELINE 10 000-point sweep itself sharded over the ranks (fixed total work); device-resident "
                          "time includes the NCCL all_gather of the eigenvalues"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs")
    dmma, dfma = engine.probe_fp64(local)
    peak = max(dmma, dfma)
    if k_real_n + k_cplx_n > 0:
        # small path: dominant kernel = s1_tridiag_kernel (assembly + Householder tridiagonalisation).
        # Algorithmic flops: (4/3) n^3 per point when the kernel takes its real-symmetric path (the
        # fluxonium H is real after the reference's 1e-12 tidy-up), (16/3) n^3 for complex Hermitian.
        real = k_real_n > 0 and k_cplx_n == 0
        kern_ms = (k_real_ms + k_cplx_ms) / args.steps
        flops_pt = (4.0 / 3.0 if real else 16.0 / 3.0) * n ** 3
        kern = ("s1t_tridiag_kernel (assemble H as 8x8 tiles of the lower triangle in shared memory + blocked Householder "
                "tridiagonalisation, trailing updates on DMMA)" if real and 80 <= n <= 128 else
                "s1_tridiag_kernel<%s> (assemble H in shared memory + Householder tridiagonalisation)" % ("double" if real else "complex"))
        note = ("FP64-pipe roofline of the dominant kernel: %s flops/point (%s path) x %d points / the kernel's device time "
                "(CUDA events on its stream, recorded by the library, averaged over warm-up + timed steps: %.3f ms of the %.3f ms step). "
                "peak = FP64 DMMA/DFMA rate measured in this run by scqed_probe_fp64_* (DMMA %.1f, DFMA %.1f TFLOP/s; "
                "MEASURED_PEAKS.json has no FP64 entry). The kernel is bound by the dependent chain of a Householder step times the "
                "matrices resident per SM (n-1 dependent steps on a %dx%d matrix, 4 CTAs per SM; profiles/r02_s1_scaling.md), not by "
                "the pipe; H never leaves shared memory; `traffic` (ncu dram bytes of this kernel, profiles/r02_s1t_tiled_ncu_summary.csv) "
                "is the lower triangle of Householder reflectors per point it writes for the back-transformation, i.e. the "
                "algorithmic output; results leaving the GPU: %.1f MB/step." % ("(4/3) n^3" if real else "(16/3) n^3", "real-symmetric" if real else "complex-Hermitian",
                                     n_pts, kern_ms, ms_per_step, dmma, dfma, n, n, d2h / 1e6))
    else:
        kern_ms = ms_per_step
        flops_pt = (16.0 / 3.0) * n ** 3
        kern = "whole step (dense / matrix-free path)"
        note = "no per-kernel span recorded; whole-step time used"
    achieved = flops_pt * n_pts / (kern_ms * 1e-3) / 1e12
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture of the same workload
    # (one launch over `TRAFFIC_PTS` points), scaled to this launch
    traffic = None
    try:
        if k_real_n + k_cplx_n > 0 and args.workload == "c2_fluxonium_res_10k":
            rows = dict((r.split(",")[0], r.strip().split(",")) for r in open(os.path.join(ROOT, "profiles", TRAFFIC_CSV)) if r.count(",") >= 2)
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            by = sum(float(rows[m][2]) * unit[rows[m][1]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            traffic = by / TRAFFIC_PTS * n_pts
    except Exception:
        traffic = None
    roofline = {
        "bound": "tensor", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": traffic, "kernel_ms_per_step": kern_ms, "kernel_share_of_step": kern_ms / ms_per_step,
        "note": note,
        "hbm_peak_gbs_measured": hbm_peak,
        "hbm_achieved_gbs_results_only": d2h / (ms_per_step * 1e-3) / 1e9,
    }
    # the HBM-bound kernel of the path, measured live beside it: dense assembly (K2, getHamiltonian for a batch)
    # of the same plan, 16 n^2 bytes written per point, CUDA events on the current stream
    k2 = None
    try:
        kp = min(8192, max(1024, n_pts))
        gk = torch.as_tensor(np.ascontiguousarray(grid[np.arange(kp) % len(grid)]), device=dev)
        Hk = torch.empty((kp, n, n), dtype=torch.complex128, device=dev)
        def k2_run():
            engine._check(eng.lib.scqed_assemble_dense(eng._handle, gk.data_ptr(), kp, 1, Hk.data_ptr(), engine._stream_ptr()))
        k2_run(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); k2_run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        k2_ms = float(np.median(ts))
        gbs = kp * 16.0 * n * n / (k2_ms * 1e-3) / 1e9
        k2 = {"kernel": "assemble_mapped_kernel / assemble_entries_kernel (scqed_assemble_dense)", "bound": "hbm",
              "points": int(kp), "ms": k2_ms, "achieved": gbs, "unit": "GB/s written", "peak": hbm_peak,
              "frac": (gbs / hbm_peak) if hbm_peak else None}
        del Hk
    except Exception as exc:                      # never let the side measurement break the bench line
        k2 = {"error": str(exc)[:200]}
    eng.close()
    del flush
    base = cpu_baseline(plan, k, vectors, resonator, budget_s=args.ref_budget)
    base_par = None
    if world == 1 and not args.no_parallel_baseline:
        base_par = cpu_baseline_parallel(args.workload, plan, k, vectors, resonator, 1.0 / max(base["value"], 1e-9))
    blocks = None
    if world == 1 and not args.no_workloads:
        blocks = []
        for name, npts, mode, nsamp, label in WORKLOADS:
            try:
                blocks.append(workload_block(name, npts, mode, nsamp, label, k, dev, peak, hbm_peak, args.workload_cpu_budget))
            except Exception as exc:              # one failing side workload must not take the headline line with it
                blocks.append({"workload": name, "error": "%s: %s" % (type(exc).__name__, str(exc)[:300])})
        try:
            blocks.extend(dense_eigh_block(dev, hbm_peak))
        except Exception as exc:
            blocks.append({"workload": "eigh_batched", "error": "%s: %s" % (type(exc).__name__, str(exc)[:300])})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (complex128 H)", "data": "synthetic",
        "config": workload_config(plan, args.workload, k, vectors, resonator, world),
        "clocks": clocks.summary(),
        "e2e": {"value": world * n_pts / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                "path": "distributed.sweep_sharded -> Engine.sweep -> scqed_sweep_run_host per rank (host parameters in, all results "
                        "out to page-locked host arrays), NCCL gather of eigenvalues / scalars / info",
                "eigenvector_dtype": ("float64 (exactly real eigenvectors: real parts packed on the device, opts.real_vectors)"
                                      if real_out else "complex128") if vectors else None,
                "pageable": {"value": world * n_pts / e2e_page_s, "ms_per_step": e2e_page_s * 1e3,
                             "note": "same call with PAGEABLE numpy result arrays: the library stages through its own pinned ring "
                                     "and a host thread moves each chunk into the caller's arrays behind the next solve"},
                "lazy_vectors": ({"value": world * n_pts / e2e_lazy_s, "ms_per_step": e2e_lazy_s * 1e3, "d2h_bytes_per_step": d2h_lazy,
                                  "note": "opts.evecs_on_device: eigenvalues + Resonator strips to the host, eigenvectors stay in HBM and "
                                          "are fetched on first use (the drop-in's default when only Resonator consumes them)"}
                                 if e2e_lazy_s else None),
                "d2h_probe": {"bytes_per_gpu": d2h, "gbs_per_gpu": d2h / probe_s / 1e9, "gbs_aggregate": world * d2h / probe_s / 1e9,
                              "ms": probe_s * 1e3,
                              "note": "all ranks copy their per-step result bytes device -> page-locked host at the same time "
                                      "(max over ranks): the PCIe / host-memory ceiling of the end-to-end figure"}},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_assembly": k2,
        "cpu_baseline": base,
        "cpu_baseline_parallel": base_par,
        "nonconverged_points": bad,
    }
    if strong is not None:
        line["strong_scaling"] = strong
    if blocks is not None:
        line["workloads"] = blocks
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_fluxonium_res_10k")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--no-vectors", action="store_true")
    ap.add_argument("--no-resonator", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=12.0, help="seconds of CPU work per reference sample")
    ap.add_argument("--complex-vectors", action="store_true", help="end-to-end leg: always ship complex128 eigenvectors")
    ap.add_argument("--no-parallel-baseline", action="store_true", help="skip the one-process-per-core CPU figure")
    ap.add_argument("--no-workloads", action="store_true", help="skip the `workloads` block (CSFQ / Averin / fluxonium-atom configs)")
    ap.add_argument("--workload-cpu-budget", type=float, default=10.0, help="seconds of CPU work per side workload")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
