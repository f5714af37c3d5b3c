"""One device-resident sweep of a golden plan, for ncu captures.
usage: python tools/prof_run.py <plan name> <points|0 = plan grid> <k> <vectors 0/1> <resonator 0/1> [solver] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycqed_b200 import engine
from pycqed_b200.plan import SweepPlan
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
name, npts, k, vec, res = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
solver = int(sys.argv[6]) if len(sys.argv) > 6 else 0
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 2
plan = SweepPlan.load(os.path.join(G, name + ".plan.json"))
grid = plan.sweep_grid()
if npts and npts < len(grid):
    grid = grid[np.unique(np.linspace(0, len(grid) - 1, npts).astype(int))]
with engine.Engine(plan) as eng:
    p = torch.as_tensor(grid, device="cuda")
    for _ in range(reps):
        out = eng.sweep_device(p, k, want_vectors=bool(vec), resonator=bool(res), solver=solver)
        torch.cuda.synchronize()
    print(name, "n", plan.hilbert_dim, "pts", len(grid), "bad", int((out[4] != 0).sum().item()))
