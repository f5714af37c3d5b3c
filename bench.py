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
same number of points (weak scaling), the only collective is the final NCCL gather of the
eigenvalues.  `--impl reference` times the CPU oracle port of the reference's per-point loop
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
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pycqed_b200 import engine

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
    grid = plan.sweep_grid()                       # every rank: the same number of points (weak scaling)
    n_pts, n = len(grid), plan.hilbert_dim
    if world > 1:
        # rank r owns the contiguous block [r*n_pts, (r+1)*n_pts) of an N-times finer global grid
        s = plan.sweep_specs[-1]
        fine = np.linspace(s["start"], s["end"], n_pts * world)
        grid = grid.copy()
        col = plan.swept_names.index(s["name"])
        if len(plan.sweep_specs) == 1:
            grid[:, col] = fine[rank * n_pts:(rank + 1) * n_pts]

    eng = engine.Engine(plan, device=local)
    p_dev = torch.as_tensor(grid, device=dev)
    p_host = torch.as_tensor(grid).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ns = plan.program.n_scalar
    nstrips = plan.post["resonator"]["nmax"] + 1 + k if resonator else 0
    # pinned host result buffers for the end-to-end leg
    hE = torch.empty((n_pts, k), dtype=torch.float64).pin_memory()
    hV = torch.empty((n_pts, k, n), dtype=torch.complex128).pin_memory() if vectors else None
    # real symmetric H (this workload): eigenvectors are exactly real and the library can ship only the real parts
    # (opts.real_vectors); a float64 pinned buffer is offered and the first call tells whether it was accepted
    hVr = torch.empty((n_pts, k, n), dtype=torch.float64).pin_memory() if (vectors and not args.complex_vectors) else None
    hR = torch.empty((n_pts, nstrips, k), dtype=torch.float64).pin_memory() if resonator else None
    out = {"E": hE.numpy(), "V": hV.numpy() if vectors else None, "Erwa": hR.numpy() if resonator else None}
    gathered = [torch.empty((n_pts, k), dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None

    def step_device():
        E, V, scal, post, info = eng.sweep_device(p_dev, k, want_vectors=vectors, resonator=resonator)
        if world > 1:
            dist.all_gather(gathered, E)
        return E, info

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
        flush.fill_(1)
    barrier()
    l0 = engine.launch_count()
    engine.timing_enable(True)
    engine.timing_read(1), engine.timing_read(2)           # clear
    times = []
    with ClockSampler(local) as clocks:
        for _ in range(args.steps):
            flush.fill_(0)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            E, info = step_device()
            e1.record()
            barrier()
            times.append(e0.elapsed_time(e1))
    engine.timing_enable(False)
    k_real_ms, k_real_n = engine.timing_read(1)
    k_cplx_ms, k_cplx_n = engine.timing_read(2)
    launches = engine.launch_count() - l0
    bad = int(info.sum().item())
    t = torch.tensor([float(np.sum(times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n_pts / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (what paramSweep does) -------------------
    real_out = False
    if hVr is not None:
        probe = eng.sweep(p_host.numpy(), k, want_vectors=True, resonator=resonator,
                          out={"E": hE.numpy(), "V": hVr.numpy(), "Erwa": hR.numpy() if resonator else None}, real_vectors="auto")
        real_out = probe.V.dtype == np.float64
        if real_out:
            out = dict(out, V=hVr.numpy())

    def step_host():
        return eng.sweep(p_host.numpy(), k, want_vectors=vectors, resonator=resonator, out=out, real_vectors=real_out)

    step_host()
    barrier()
    e2e_times = []
    for _ in range(max(1, args.steps)):
        flush.fill_(0)
        barrier()
        t0 = time.perf_counter()
        r = step_host()
        if world > 1:
            dist.all_gather(gathered, torch.as_tensor(r.E, device=dev))
            torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
    te = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d = int(grid.nbytes)
    d2h = int(hE.numel() * 8 + (hV.numel() * (8 if real_out else 16) if vectors else 0) + (hR.numel() * 8 if resonator else 0)
              + n_pts * max(ns, 1) * 8 + n_pts * 4)

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
    dmma, dfma = engine.probe_fp64(local)
    peak = max(dmma, dfma)
    if k_real_n + k_cplx_n > 0:
        # small path: dominant kernel = s1_tridiag_kernel (assembly + Householder tridiagonalisation).
        # Algorithmic flops: (4/3) n^3 per point when the kernel takes its real-symmetric path (the
        # fluxonium H is real after the reference's 1e-12 tidy-up), (16/3) n^3 for complex Hermitian.
        real = k_real_n > 0 and k_cplx_n == 0
        kern_ms = (k_real_ms + k_cplx_ms) / args.steps
        flops_pt = (4.0 / 3.0 if real else 16.0 / 3.0) * n ** 3
        kern = "s1_tridiag_kernel<%s> (assemble H in shared memory + Householder tridiagonalisation)" % ("double" if real else "complex")
        note = ("FP64-pipe roofline of the dominant kernel: %s flops/point (%s path) x %d points / the kernel's device time "
                "(CUDA events on its stream, recorded by the library, averaged over the timed steps: %.3f ms of the %.3f ms step). "
                "peak = FP64 DMMA/DFMA rate measured in this run by scqed_probe_fp64_* (DMMA %.1f, DFMA %.1f TFLOP/s; "
                "MEASURED_PEAKS.json has no FP64 entry). The kernel is latency/barrier bound (n-1 dependent Householder steps on "
                "a 101x101 matrix), not pipe bound; H never leaves shared memory; `traffic` (ncu dram bytes of this kernel) is the "
                "n^2 x 8 B of Householder reflectors per point it writes for the back-transformation, i.e. the algorithmic output; "
                "results leaving the GPU: %.1f MB/step." % ("(4/3) n^3" if real else "(16/3) n^3", "real-symmetric" if real else "complex-Hermitian",
                                     n_pts, kern_ms, ms_per_step, dmma, dfma, d2h / 1e6))
    else:
        kern_ms = ms_per_step
        flops_pt = (16.0 / 3.0) * n ** 3
        kern = "whole step (dense / matrix-free path)"
        note = "no per-kernel span recorded; whole-step time used"
    achieved = flops_pt * n_pts / (kern_ms * 1e-3) / 1e12
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture of the same workload
    # (profiles/r01_s1_tridiag_ncu_summary.csv: one launch over 9704 points), scaled to this launch
    traffic = None
    try:
        if k_real_n + k_cplx_n > 0 and args.workload == "c2_fluxonium_res_10k":
            rows = dict((r.split(",")[0], r.split(",")) for r in open(os.path.join(ROOT, "profiles", "r01_s1_tridiag_ncu_summary.csv")) if r.count(",") >= 2)
            mb = float(rows["dram__bytes_read.sum"][2]) + float(rows["dram__bytes_write.sum"][2])
            traffic = mb * 1e6 / 9704.0 * n_pts
    except Exception:
        traffic = None
    roofline = {
        "bound": "tensor", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": traffic, "kernel_ms_per_step": kern_ms, "kernel_share_of_step": kern_ms / ms_per_step,
        "note": note,
        "hbm_peak_gbs_measured": peaks.get("hbm_gbs"),
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
              "points": int(kp), "ms": k2_ms, "achieved": gbs, "unit": "GB/s written", "peak": peaks.get("hbm_gbs"),
              "frac": (gbs / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None}
        del Hk
    except Exception as exc:                      # never let the side measurement break the bench line
        k2 = {"error": str(exc)[:200]}
    base = cpu_baseline(plan, k, vectors, resonator, budget_s=args.ref_budget)
    base_par = None
    if world == 1 and not args.no_parallel_baseline:
        base_par = cpu_baseline_parallel(args.workload, plan, k, vectors, resonator, 1.0 / max(base["value"], 1e-9))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (complex128 H)", "data": "synthetic",
        "config": workload_config(plan, args.workload, k, vectors, resonator, world),
        "clocks": clocks.summary(),
        "e2e": {"value": world * n_pts / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3,
                "eigenvector_dtype": ("float64 (exactly real eigenvectors: real parts packed on the device, opts.real_vectors)"
                                      if real_out else "complex128") if vectors else None},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_assembly": k2,
        "cpu_baseline": base,
        "cpu_baseline_parallel": base_par,
        "nonconverged_points": bad,
    }
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
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
