"""One K2 (scqed_assemble_dense) call for ncu captures: python tools/prof_k2.py <plan> <points>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycqed_b200 import engine
from pycqed_b200.plan import SweepPlan
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
plan = SweepPlan.load(os.path.join(G, sys.argv[1] + ".plan.json"))
pts = int(sys.argv[2])
grid = plan.sweep_grid()
grid = grid[np.arange(pts) % len(grid)]
with engine.Engine(plan) as eng:
    for _ in range(2):
        H = eng.assemble_dense(grid)
        torch.cuda.synchronize()
print("ok", H.shape)
