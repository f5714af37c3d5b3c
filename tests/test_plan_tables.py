"""Host-side plan tables (compile-time work on a SweepPlan) against the oracle's Hamiltonian:
the dense real term table of the fused small-n solver (pycqed_b200/densetab.py) and the entry list of
the dense assembly kernel (pycqed_b200/entrytab.py).  Both must reproduce
H(p) = sum_t c_t(p) kron_k F_{t,k}  (pyscqed/numerical_system.py:509-594) exactly as the term table does."""
import numpy as np
import pytest

from oracle.plan_oracle import PlanOracle, _run_program, eval_program
from tests.conftest import load_golden, load_plan

SMALL = ["c1_transmon", "c2_fluxonium_res", "c4_averin", "c6_rfsquid", "c6_cpb", "c6_csfq3jj_t5", "c7_fluxonium_LphiZ",
         "c7_fluxonium_C", "c2_fluxonium_t200s"]


@pytest.mark.parametrize("name", SMALL)
def test_dense_term_table_reproduces_hamiltonian(name):
    from pycqed_b200.densetab import build_dense_table
    g, plan = load_golden(name)
    dt = build_dense_table(plan)
    assert dt is not None
    n = plan.hilbert_dim
    nf, nd = dt["n_full"], dt["n_diag"]
    assert len(dt["re_regs"]) == nf + nd == len(dt["im_regs"]) == len(dt["gmax"])
    assert dt["table"].size == nf * n * n + nd * n
    # the appended instructions leave the original outputs untouched
    prm = g["params"][:: max(1, len(g["params"]) // 5)]
    c0, s0 = eval_program(plan.program, prm)
    c1, s1 = eval_program(dt["program"], prm)
    assert np.array_equal(c0, c1) and np.array_equal(s0, s1)
    regs = _run_program(dt["program"], prm)
    orc = PlanOracle(plan)
    full = dt["table"][: nf * n * n].reshape(nf, n, n).transpose(0, 2, 1)      # [g][i][j]
    diag = dt["table"][nf * n * n:].reshape(nd, n)
    tri = True
    for p in range(len(prm)):
        H = orc.hamiltonian(c0[p], tidy=False).toarray()
        A = np.zeros((n, n), dtype=np.complex128)
        for gi in range(nf + nd):
            al = regs[dt["re_regs"][gi], p] if dt["re_regs"][gi] >= 0 else 0.0
            be = regs[dt["im_regs"][gi], p] if dt["im_regs"][gi] >= 0 else 0.0
            A += (al + 1j * be) * (full[gi] if gi < nf else np.diag(diag[gi - nf]))
        assert np.max(np.abs(A - H)) <= 1e-13 * max(1.0, np.abs(H).max())
        tri &= not np.any(np.triu(H, 2)) and not np.any(np.tril(H, -2))
    if dt["tridiagonal"]:
        assert tri, "flagged tridiagonal but H has entries outside the three diagonals"
    if name in ("c1_transmon", "c4_averin", "c6_cpb"):
        assert dt["tridiagonal"]            # single-node charge basis
    if name.startswith("c2_") or name.startswith("c6_rfsquid"):
        assert nf <= 3                      # constant part + Re D + Im D


def test_dense_table_limits():
    from pycqed_b200.densetab import MAX_N, build_dense_table
    assert build_dense_table(load_plan("c3_csfq4jj_t10")) is None        # n = 441 > MAX_N
    assert MAX_N >= 212


@pytest.mark.parametrize("name", ["c1_transmon", "c2_fluxonium_res", "c3_csfq4jj_t10", "c3_csfqfloat_t5",
                                  "c5_fluxatom_t31", "c7_fluxonium_C"])
def test_entry_list_reproduces_hamiltonian(name):
    from pycqed_b200.entrytab import build_entry_table
    g, plan = load_golden(name)
    et = build_entry_table(plan)
    assert et is not None
    n = plan.hilbert_dim
    assert np.all(np.diff(et["key"]) > 0) and et["key"][0] >= 0 and et["key"][-1] < n * n
    assert et["ptr"][0] == 0 and et["ptr"][-1] == len(et["term"]) == len(et["val"]) and np.all(np.diff(et["ptr"]) > 0)
    coef, _ = eval_program(plan.program, g["params"][:3])
    orc = PlanOracle(plan)
    for p in range(len(coef)):
        H = orc.hamiltonian(coef[p], tidy=False).toarray()
        A = np.zeros(n * n, dtype=np.complex128)
        A[et["key"]] = np.add.reduceat(coef[p][et["term"]] * et["val"], et["ptr"][:-1])
        assert np.max(np.abs(A.reshape(n, n) - H)) <= 1e-13 * max(1.0, np.abs(H).max())
    assert et["dense"] == (4 * len(et["key"]) >= n * n)


def test_entry_list_size_guard():
    from pycqed_b200.entrytab import build_entry_table
    assert build_entry_table(load_plan("c5_fluxatom_t180"), max_pairs=1_000_000) is None


def test_regen_operator_set_is_a_similarity():
    """operators.oscillator_regen_set: W^T (a + a^dag) W is diagonal, the charge operator keeps its spectrum, and
    exp(i beta X) = W diag(exp(i beta lam)) W^T is the reference's D (numerical_system.py:1078-1090)."""
    from pycqed_b200 import operators as ops
    d, z = 24, 137.0
    rs = ops.oscillator_regen_set(d)
    a = ops.oscillator_ladder(d)
    X = (a + a.conj().T).real
    assert np.max(np.abs(rs["W"].T @ X @ rs["W"] - np.diag(rs["lam"]))) < 1e-12
    assert np.max(np.abs(rs["W"].T @ rs["W"] - np.eye(d))) < 1e-13
    Q, P, D, Dd, S, Sd = ops.oscillator_basis(d, z, 1.0, 1.0, 1.0, 1.0)
    beta, _ = ops.oscillator_displacement_amplitudes(z, 1.0, 1.0)
    Dw = rs["W"] @ np.diag(np.exp(1j * beta * rs["lam"])) @ rs["W"].T
    assert np.max(np.abs(Dw - D)) < 1e-11
    q = np.sqrt(1.0 / (2.0 * z))
    assert np.max(np.abs(q * rs["W"] @ rs["charge"] @ rs["W"].T - Q)) < 1e-12
