#!/usr/bin/env python
"""A/B of the S1 occupancy ladder (SCQED_S1_LADDER) on the bench workload: tridiagonalisation device time (library
CUDA events, tag 1) and whole device-resident step.   python tools/s1_ladder.py [setting ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycqed_b200 import engine
from pycqed_b200.plan import SweepPlan
plan = SweepPlan.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "c2_fluxonium_res_10k.plan.json"))
g = torch.as_tensor(plan.sweep_grid(), device="cuda")
settings = sys.argv[1:] or ["0", "auto", "72,48", "72", "0"]
ref = None
with engine.Engine(plan) as eng:
    for s in settings:
        os.environ["SCQED_S1_TILED"] = "0"
        os.environ.pop("SCQED_S1_TILED_ROTATE", None)
        os.environ.pop("SCQED_S1_TILED_PSER", None)
        if s == "tiled-1warp":
            os.environ["SCQED_S1_TILED_PSER"] = "0"
        if s.startswith("tiled"):
            os.environ["SCQED_S1_TILED"] = "1"
            if s == "tiled-rot":
                os.environ["SCQED_S1_TILED_ROTATE"] = "1"
            if s == "tiled-ded":
                os.environ["SCQED_S1_TILED_ROTATE"] = "2"
        elif s == "auto":
            os.environ.pop("SCQED_S1_LADDER", None)
        else:
            os.environ["SCQED_S1_LADDER"] = s
        for _ in range(3):
            out = eng.sweep_device(g, 10, want_vectors=True, resonator=True)
        torch.cuda.synchronize()
        engine.timing_enable(True); engine.timing_read(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = eng.sweep_device(g, 10, want_vectors=True, resonator=True)
        e1.record(); torch.cuda.synchronize()
        ms, c = engine.timing_read(1); engine.timing_enable(False)
        E = out[0].cpu().numpy()
        if ref is None:
            ref = E
        print("ladder=%-14s tridiag %.3f ms  step %.3f ms  bad=%d  max|dE| vs first %.2e" % (s, ms / 5, e0.elapsed_time(e1) / 5, int(out[4].sum()), np.abs(E - ref).max()), flush=True)
