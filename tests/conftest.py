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


def golden_vectors(g, a):
    """Eigenvectors [n, k] of stored sample ``a`` as complex128.  The big fixtures keep them as complex64; rounding
    changes the NORM at first order (the direction only at second order), so columns are re-normalised."""
    V = np.asarray(g["V"][a])
    if V.dtype != np.complex128:
        V = V.astype(np.complex128)
        V = V / np.linalg.norm(V, axis=0, keepdims=True)
    return V


_NEXT_CACHE = {}


def e_next(g, plan, point, k):
    """Level k (the first one the k-truncation drops) at fixture point ``point``: stored by the reference run
    when the fixture was minted with ``k_extra`` (``E_next``), otherwise solved by the oracle.  It tells
    ``subspace_overlap`` whether level k-1 sits in a degenerate cluster that the truncation cuts."""
    if "E_next" in g:
        return float(g["E_next"][point][0])
    n = plan.hilbert_dim
    if k >= n:
        return None
    key = (plan.meta.get("source"), int(point), int(k), tuple(np.asarray(g["params"][point]).tolist()))
    if key not in _NEXT_CACHE:
        from oracle.plan_oracle import PlanOracle, eval_program
        orc = _NEXT_CACHE.setdefault(("oracle", id(plan)), PlanOracle(plan))
        coef, _ = eval_program(plan.program, np.atleast_2d(g["params"][point]))
        H = orc.hamiltonian(coef[0])
        if n <= 2000:
            E = orc.diag_dense(H, k + 1)
        else:
            E = orc.diag_sparse(H, k + 1, which="SA", tol=0)
        _NEXT_CACHE[key] = float(np.sort(E)[k])
    return _NEXT_CACHE[key]
