import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("PYSCQED_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference checkout at /root/reference")


def has_reference():
    return os.path.isdir(os.path.join(REFERENCE, "pyscqed"))


needs_reference = pytest.mark.skipif(not has_reference(), reason="reference checkout not present")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name):
    from pycqed_b200.plan import SweepPlan
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    plan = SweepPlan.load(os.path.join(GOLDEN, name + ".plan.json"))
    return g, plan


def load_plan(name):
    from pycqed_b200.plan import SweepPlan
    return SweepPlan.load(os.path.join(GOLDEN, name + ".plan.json"))


def golden_hamiltonian(g, j, n):
    import scipy.sparse as sp
    return sp.csr_matrix((g["H%d_val" % j], (g["H%d_row" % j], g["H%d_col" % j])), shape=(n, n))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
