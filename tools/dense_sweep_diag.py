"""Where a plan-based dense real symmetric sweep (fluxonium at a large node truncation) spends its time:
device-resident vs host call, reduction span (tag 8), launch count.   python tools/dense_sweep_diag.py [trunc] [points]"""
import sys, json, time, os
sys.path.insert(0, ".")
import numpy as np, torch
from pycqed_b200 import engine
from pycqed_b200.plan import SweepPlan
trunc = int(sys.argv[1]) if len(sys.argv) > 1 else 640
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 592
d = json.load(open("tests/golden/c2_fluxonium_t200.plan.json")); d["nodes"][0]["trunc"] = trunc
plan = SweepPlan.from_dict(d)
full = plan.sweep_grid()
grid = np.ascontiguousarray(full[np.linspace(0, len(full) - 1, npts).round().astype(int)])
with engine.Engine(plan) as eng:
    for rep in range(3):
        engine.timing_enable(True); engine.timing_read(8)
        l0 = engine.launch_count()
        t0 = time.perf_counter(); res = eng.sweep(grid, 10, want_vectors=True); t1 = time.perf_counter()
        red, cnt = engine.timing_read(8); engine.timing_enable(False)
        print("host sweep %.1f ms  (reduction spans %.1f ms in %d spans, %d launches)" % ((t1 - t0) * 1e3, red, cnt, engine.launch_count() - l0), flush=True)
    p_dev = torch.as_tensor(grid, device="cuda")
    for rep in range(3):
        torch.cuda.synchronize()
        engine.timing_enable(True); engine.timing_read(8)
        t0 = time.perf_counter(); out = eng.sweep_device(p_dev, 10, want_vectors=True); torch.cuda.synchronize(); t1 = time.perf_counter()
        red, cnt = engine.timing_read(8); engine.timing_enable(False)
        print("device sweep %.1f ms  (reduction spans %.1f ms in %d spans)" % ((t1 - t0) * 1e3, red, cnt), flush=True)
