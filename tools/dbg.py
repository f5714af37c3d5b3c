import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from pycqed_b200 import engine
from tools.diag import load
for name in ["c1_transmon", "c4_averin"]:
    g, plan = load(name)
    with engine.Engine(plan) as eng:
        res = eng.sweep(g["params"], int(g["k"]))
    bad = np.nonzero(res.info)[0]
    print(name, "bad idx", bad[:20], "n bad", len(bad))
    for i in bad[:4]:
        print("  params", g["params"][i], "\n   E  ", res.E[i], "\n   ref", g["E"][i])
