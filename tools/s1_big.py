import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from pycqed_b200 import engine
from pycqed_b200.plan import SweepPlan
name = sys.argv[1] if len(sys.argv) > 1 else "c2_fluxonium_t200"
plan = SweepPlan.load("tests/golden/%s.plan.json" % name)
g = torch.as_tensor(plan.sweep_grid(), device="cuda")
ref = None
for big in ("0", "1"):
    os.environ["SCQED_S1_TILED_BIG"] = big
    with engine.Engine(plan) as eng:
        for _ in range(2): out = eng.sweep_device(g, 10, want_vectors=True)
        torch.cuda.synchronize(); engine.timing_enable(True); engine.timing_read(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): out = eng.sweep_device(g, 10, want_vectors=True)
        e1.record(); torch.cuda.synchronize()
        ms, c = engine.timing_read(1); engine.timing_enable(False)
    E = out[0].cpu().numpy()
    if ref is None: ref = E
    print("%s n=%d BIG=%s tridiag %.2f ms step %.2f ms (%.0f pts/s) bad=%d dE %.2e" % (name, plan.hilbert_dim, big, ms / 3, e0.elapsed_time(e1) / 3, len(g) / (e0.elapsed_time(e1) / 3e3), int(out[4].sum()), np.abs(E - ref).max()))
