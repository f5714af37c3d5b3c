"""GPU tests of what sits between the user and the kernels: the drop-in class itself, the single-point
(zero swept columns) path, the multi-engine sweep driver, the host-buffer staging of ``scqed_sweep_run_host``
and the real -> complex fall-back of the shared-memory solver.

The GPU box has no reference checkout, so ``NumericalSystem`` is driven with a stub symbolic system that answers
the handful of ``SymbolicSystem`` calls the class makes, and ``lower_system`` is patched to hand back the plans
that tests/golden/make_golden.py lowered from the real reference objects (same requests the class makes).
Everything numerical runs on the device and is compared with outputs of the unmodified reference (goldens).
"""
import numpy as np
import pytest
import sympy as sy

from oracle.plan_oracle import PlanOracle, eig_rel_err, eval_program, subspace_overlap
from tests.conftest import e_next, golden_hamiltonian, load_golden, load_plan

pytestmark = pytest.mark.gpu

E_RTOL = 1e-9
OVERLAP = 1 - 1e-8


@pytest.fixture(scope="module")
def engine_mod(cuda_device):
    from pycqed_b200 import engine
    engine.load_library()
    return engine


# ---------------------------------------------------------------------------------------------
# stub symbolic layer
# ---------------------------------------------------------------------------------------------
class _Units:
    def getUnitPrefactor(self, name):
        return 1.0

    def getPrefactor(self, name):
        return 1.0


class _StubSS:
    """The slice of pyscqed.SymbolicSystem / ParamCollection the drop-in class touches (method names as in
    symbolic_system.py / parameters.py)."""

    def __init__(self, nodes, edges, values):
        self.nodes, self.edges = list(nodes), list(edges)
        self._values = dict(values)
        self.sweeps = None

    def _has_resonators(self):
        return False

    def getInverseCapacitanceMatrix(self):
        return sy.Matrix(len(self.nodes), len(self.nodes), lambda i, j: sy.Symbol("C%d%di" % (min(i, j), max(i, j))))

    def getInverseInductanceMatrix(self):
        return sy.Matrix(len(self.nodes), len(self.nodes), lambda i, j: sy.Symbol("L%d%di" % (min(i, j), max(i, j))))

    def getJosephsonVector(self):
        return sy.Matrix([sy.Symbol("J%d" % i) for i in range(len(self.edges))])

    def getPhaseSlipVector(self):
        return sy.Matrix([sy.Symbol("V%d" % i) for i in range(len(self.edges))])

    def getParameterValuesDict(self):
        return dict(self._values)

    def getParametricParametersList(self):
        return []

    def getParametricExpression(self, name):
        raise KeyError(name)

    def getParameterNamesList(self):
        return list(self._values)

    def setParameterValue(self, name, value):
        self._values[name] = value

    def getParameterValue(self, name):
        return self._values[name]

    def paramSweepSpec(self, name, start, end, N):
        return {"name": name, "start": start, "end": end, "N": N}

    def ndSweep(self, specs):
        self.sweeps = specs


def _system(monkeypatch, plans, nodes, edges, values, basis, trunc, device=0):
    """A drop-in NumericalSystem whose lowering requests are answered from stored plans:
    ``plans[(swept names, resonator requested, matrix elements requested)]``."""
    import pycqed_b200.numerical_system as ns

    def fake_lower(SS, units, operator_data, swept_names=(), sweep_specs=(), scalars=None, resonator=None,
                   matrix_elements=(), regen_mode="consistent", matrices=None):
        return plans[(tuple(swept_names), resonator is not None, bool(matrix_elements))]
    monkeypatch.setattr(ns, "lower_system", fake_lower)
    h = ns.NumericalSystem(_StubSS(nodes, edges, values), unit=_Units(), device=device)
    for nd in nodes:
        h.operator_data[nd] = {"truncation": trunc, "basis": basis, "impedance": None, "frequency": None, "flux_max": None}
    return h, ns


# ---------------------------------------------------------------------------------------------
# the class on the GPU
# ---------------------------------------------------------------------------------------------
def test_numerical_system_single_point_api(engine_mod, monkeypatch):
    """getHamiltonian / eigenenergies / diagonalize / getResonatorResponse / get*Energies at the current parameter
    values (numerical_system.py:509,790-810,736-784,696-734): plans with ZERO swept columns, params of shape (1, 0)."""
    g, plan = load_golden("c8_single_fluxonium")
    plans = {((), False, False): plan, ((), True, False): load_plan("c8_single_fluxonium_res")}
    h, ns = _system(monkeypatch, plans, [1], [(0, 1, 1), (1, 0, 2)], {"phiZ": 0.5}, "oscillator", 101)
    n, k = int(g["n"]), int(g["k"])
    H = h.getHamiltonian()
    Href = golden_hamiltonian(g, 0, n).toarray()
    assert H.shape == (n, n) and H.dims == [[n], [n]]
    assert np.max(np.abs(H.full() - Href)) <= 4e-12 * max(1.0, np.abs(Href).max())
    assert H.isherm
    w = np.linalg.eigvalsh(Href)
    assert np.max(np.abs(H.eigenenergies() - w)) < 1e-9 * np.abs(w).max()
    h.setDiagConfig(eigvalues=k, get_vectors=True)
    E, V = h.diagonalize(H)
    assert isinstance(E, tuple) and len(E) == k and V.shape == (k,) and V.dtype == object
    assert eig_rel_err(np.asarray(E), g["E"][0]) < E_RTOL
    X = np.stack([v.full()[:, 0] for v in V], axis=1)
    assert V[0].dims == [[n], [1]] and abs(V[3].norm() - 1.0) < 1e-12
    assert subspace_overlap(X, g["V"][0], g["E"][0], E_next=w[k]) >= OVERLAP
    h.setDiagConfig(eigvalues=k, get_vectors=False)
    E2 = h.diagonalize(H)
    assert isinstance(E2, np.ndarray) and eig_rel_err(E2, g["E"][0]) < E_RTOL
    Erwa = h.getResonatorResponse(np.asarray(E), V, nmax=100, cpl_node=1)
    ref = g["eval_getResonatorResponse"][0]
    assert Erwa.shape == ref.shape and np.max(np.abs(Erwa - ref)) < E_RTOL * np.abs(ref).max()
    assert h.getResonatorResponse(np.asarray(E), V) is None                       # :737-741
    # scalar evaluables: every node / edge at once (dict) and one by one
    names = [str(x) for x in g["scalar_names"]]
    vals = dict(zip(names, g["scalar_values"]))
    ce = h.getChargingEnergies()
    assert set(ce) == {1} and abs(ce[1] - vals["ChargingEnergy:1"]) <= 1e-13 * abs(vals["ChargingEnergy:1"])
    assert abs(h.getFluxEnergies(1) - vals["FluxEnergy:1"]) <= 1e-13 * abs(vals["FluxEnergy:1"])
    je = h.getJosephsonEnergies()
    for edge, v in je.items():
        assert abs(v - vals["JosephsonEnergy:%r" % (edge,)]) <= 1e-13 * max(1.0, abs(v))
    with pytest.raises(Exception):
        h.diagonalize(np.eye(3))                                                  # util.py:96-97: not a Qobj
    h.close()


def test_numerical_system_two_node_single_point(engine_mod, monkeypatch):
    g, plan = load_golden("c8_single_csfq3jj")
    h, ns = _system(monkeypatch, {((), False, False): plan}, [1, 2], [(0, 1, 0), (0, 2, 0), (2, 1, 0)], {"phiZ": 0.5}, "charge", 6)
    n, k = int(g["n"]), int(g["k"])
    H = h.getHamiltonian()
    Href = golden_hamiltonian(g, 0, n).toarray()
    assert H.dims == [[13, 13], [13, 13]]
    assert np.max(np.abs(H.full() - Href)) <= 4e-12 * max(1.0, np.abs(Href).max())
    h.setDiagConfig(eigvalues=k, get_vectors=True)
    E, V = h.diagonalize(H)
    assert eig_rel_err(np.asarray(E), g["E"][0]) < E_RTOL
    assert V[0].dims == [[13, 13], [1, 1]]
    X = np.stack([v.full()[:, 0] for v in V], axis=1)
    assert subspace_overlap(X, g["V"][0], g["E"][0], E_next=np.linalg.eigvalsh(Href)[k]) >= OVERLAP
    h.close()


@pytest.mark.parametrize("lazy", [True, False])
def test_numerical_system_param_sweep_on_gpu(engine_mod, monkeypatch, lazy):
    """paramSweep -> EnginePool -> scqed_sweep_run_host -> SweepData -> getSweep on CUDA, compared with the
    reference's own paramSweep output in the layouts getSweepResult gives: (k, N), (k, 2, N), (k, 121, N)."""
    g, plan = load_golden("c2_fluxonium_res")
    h, ns = _system(monkeypatch, {(("phiZ",), True, False): plan, (("phiZ",), False, False): load_plan("c2_fluxonium")},
                    [1], [(0, 1, 1), (1, 0, 2)], {"phiZ": 0.5}, "oscillator", 101)
    h.lazy_vectors = lazy
    k, N = int(g["k"]), len(g["E"])
    h.newSweep()
    h.addSweep("phiZ", -1.0, 1.0, N)
    h.addEvaluation("Hamiltonian")
    h.addEvaluation("Resonator", cpl_node=1)
    h.setDiagConfig(eigvalues=k, get_vectors=True)
    data = h.paramSweep(timesweep=False)
    assert h.sweep_specs == [] and h.evaluations == []                              # consumed (:915)
    assert len(data) == N and (data._res.V is None) == lazy
    x, EV, sv = h.getSweep(data, "phiZ", {})
    assert np.allclose(x, g["params"][:, 0]) and sv == {}
    assert EV.shape == (k, 2, N)
    E, V = EV[:, 0], EV[:, 1]
    assert E.shape == (k, N) and V.shape == (k, N)
    assert eig_rel_err(E.T, g["E"]) < E_RTOL
    for a, p in enumerate(g["vec_points"]):
        X = np.stack([V[j, p].full()[:, 0] for j in range(k)], axis=1)
        assert subspace_overlap(X, g["V"][a], g["E"][p], E_next=e_next(g, plan, p, k)) >= OVERLAP
    _, R, _ = h.getSweep(data, "phiZ", {}, evaluable="Resonator")
    ref = g["eval_getResonatorResponse"]                                            # [N, 121, k]
    assert R.shape == (k, ref.shape[1], N)
    assert np.max(np.abs(R - ref.T)) < E_RTOL * np.abs(ref).max()
    # per-point records in the reference's format
    rec = data.record(3)
    assert set(rec) == {"getHamiltonian", "getResonatorResponse"}
    assert isinstance(rec["getHamiltonian"][0], tuple) and rec["getHamiltonian"][1].shape == (k,)
    # second sweep of the same circuit: plan and engine come from the cache
    h.addSweep("phiZ", -0.3, 0.3, 7)
    h.addEvaluation("Hamiltonian")
    h.addEvaluation("Resonator", cpl_node=1)
    d2 = h.paramSweep(timesweep=False)
    assert d2.timings["init"] < 0.002 and d2.E.shape == (7, k) and d2.Erwa.shape == (7, 100 + 1 + k, k)
    # a different request (no Resonator): lowered / uploaded once more, then cached as well
    h.addSweep("phiZ", -0.3, 0.3, 7)
    h.setDiagConfig(eigvalues=5)
    d3 = h.paramSweep(timesweep=False)
    assert d3.E.shape == (7, 5) and eig_rel_err(d3.E, d2.E[:, :5]) < 1e-12
    h.close()


def test_zero_swept_columns_through_every_entry_point(engine_mod):
    """Plans without a swept parameter: params of shape (n_pts, 0) through K1, K2 and the sweep (device and host)."""
    g, plan = load_golden("c8_single_fluxonium")
    gr, plan_r = load_golden("c8_single_fluxonium")[0], load_plan("c8_single_fluxonium_res")
    n, k = int(g["n"]), int(g["k"])
    p0 = np.zeros((1, 0))
    with engine_mod.Engine(plan) as eng:
        coef, scal = eng.coefficients(p0)
        cref, sref = eval_program(plan.program, p0)
        assert np.max(np.abs(coef.cpu().numpy() - cref)) <= 1e-14 * np.abs(cref).max()
        assert np.max(np.abs(scal.cpu().numpy() - sref)) <= 1e-13 * np.abs(sref).max()
        H = eng.assemble_dense(p0).cpu().numpy()[0]
        Href = golden_hamiltonian(g, 0, n).toarray()
        assert np.max(np.abs(H - Href)) <= 4e-12 * max(1.0, np.abs(Href).max())
        for solver in (0, 2):
            res = eng.sweep(p0, k, want_vectors=True, solver=solver)
            assert not res.info.any() and eig_rel_err(res.E, g["E"]) < E_RTOL
        E, V, _, _, info = eng.sweep_device(np.zeros((3, 0)), k, want_vectors=True)      # three copies of the point
        assert np.array_equal(E.cpu().numpy()[0], E.cpu().numpy()[2]) and eig_rel_err(E.cpu().numpy(), np.repeat(g["E"], 3, 0)) < E_RTOL
    with engine_mod.Engine(plan_r) as eng:
        res = eng.sweep(p0, k, want_vectors=True, resonator=True)
        ref = gr["eval_getResonatorResponse"]
        assert np.max(np.abs(res.Erwa - ref)) < E_RTOL * np.abs(ref).max()


# ---------------------------------------------------------------------------------------------
# sweep driver: several engines, staging, lazy vectors, argument checking
# ---------------------------------------------------------------------------------------------
def test_engine_pool_shards_like_one_engine(engine_mod):
    """EnginePool = one host thread per engine on contiguous blocks of the grid, all writing rows of the same pinned
    arrays.  Two engines on device 0 exercise the threaded path on a one-GPU box; with more GPUs visible the same
    test also runs across devices.  Results equal the single-engine sweep bit for bit."""
    import torch
    plan = load_plan("c2_fluxonium_res_10k")
    grid = plan.sweep_grid()[:3000]
    with engine_mod.Engine(plan) as eng:
        one = eng.sweep(grid, 10, want_vectors=True, resonator=True)
    devs = [0, 0] if torch.cuda.device_count() < 2 else list(range(torch.cuda.device_count()))
    with engine_mod.EnginePool(plan, devs) as pool:
        many = pool.sweep(grid, 10, want_vectors=True, resonator=True, real_vectors="auto")
        lazy = pool.sweep(grid, 10, want_vectors=True, resonator=True, vectors_on_device=True)
    assert not one.info.any() and not many.info.any()
    # blocks of >= 410 points take the same kernels as the whole sweep (bit-identical results); smaller blocks (8 GPUs:
    # 375 points each) use the warp multisection instead of the thread bisection: same tolerance, last-ulp differences
    same_path = len(grid) // len(devs) * 10 >= 4096
    scale = np.abs(one.E).max()

    def close(a, b, tol):
        return np.array_equal(a, b) if same_path else np.max(np.abs(a - b)) <= tol * max(1.0, np.abs(a).max())
    assert close(one.E, many.E, 1e-13) and close(one.Erwa, many.Erwa, 1e-11), scale
    assert many.V.dtype == np.float64 and close(one.V.real, many.V, 1e-9)
    assert lazy.V is None and len(lazy.V_dev) == len(devs)
    assert close(lazy.vectors(), one.V, 1e-9) and close(lazy.Erwa, one.Erwa, 1e-11)
    for nm in one.scalars:
        assert np.array_equal(one.scalars[nm], many.scalars[nm])


def test_pageable_and_pinned_host_buffers_agree(engine_mod, monkeypatch):
    """scqed_sweep_run_host: pageable caller arrays go through the library's pinned staging ring, pinned ones are
    written directly; SCQED_FORCE_STAGING stages everything.  Same bytes either way, complex and real vectors."""
    plan = load_plan("c2_fluxonium_res_10k")
    grid = plan.sweep_grid()[:6000]          # > 5000 points: the call is cut into chunks
    k, n = 10, plan.hilbert_dim
    ns_ = 100 + 1 + k
    with engine_mod.Engine(plan) as eng:
        a = eng.sweep(grid, k, want_vectors=True, resonator=True)                                  # pinned (library default)
        out = {"E": np.empty((6000, k)), "V": np.empty((6000, k, n), dtype=np.complex128), "Erwa": np.empty((6000, ns_, k)),
               "info": np.empty(6000, dtype=np.int32)}
        b = eng.sweep(grid, k, want_vectors=True, resonator=True, out=out)                         # pageable numpy arrays
        monkeypatch.setenv("SCQED_FORCE_STAGING", "1")
        c = eng.sweep(grid, k, want_vectors=True, resonator=True, real_vectors=True)
        monkeypatch.delenv("SCQED_FORCE_STAGING")
        assert np.shares_memory(b.E, out["E"]) and np.shares_memory(b.V, out["V"])
        for r in (b, c):
            assert not r.info.any()
            assert np.array_equal(a.E, r.E) and np.array_equal(a.Erwa, r.Erwa)
        assert np.array_equal(a.V, b.V) and np.array_equal(a.V.real, c.V)
        # argument checking of caller buffers (raw pointers cross the ABI)
        for bad in ({"E": np.empty((6000, k + 1))}, {"E": np.empty((6000, k), dtype=np.float32)},
                    {"V": np.empty((6000, n, k), dtype=np.complex128)}, {"Erwa": np.empty((6000, ns_, 2 * k))[:, :, ::2]},
                    {"E": np.empty((100, k))}):
            with pytest.raises(engine_mod.EngineError):
                eng.sweep(grid, k, want_vectors=True, resonator=True, out=bad)


def test_real_plan_turning_complex_late_in_the_grid(engine_mod):
    """A plan probed real symmetric on its first 2 x SM-count points and complex afterwards: the real kernel flags the
    complex points, the device path counts the flags and redoes the call in complex mode -- through sweep_device,
    sweep(matrix_elements=True) and the host path alike (the first two used to return stale workspaces)."""
    g, plan = load_golden("c6_csfq3jj_t5")
    k = int(g["k"])
    col = plan.swept_names.index("phiZ")
    params = np.zeros((700, len(plan.swept_names)))
    params[:, col] = np.concatenate([np.zeros(600), np.linspace(0.07, 0.43, 100)])        # phiZ = 0: H is real
    orc = PlanOracle(plan)
    idx = np.array([0, 299, 599, 600, 650, 699])
    ref = orc.sweep(params[idx], k + 1, get_vectors=True)
    for mode in ("device", "matel", "host"):
        with engine_mod.Engine(plan) as eng:
            if mode == "device":
                E, V, _, _, info = eng.sweep_device(params, k, want_vectors=True)
                E, V, info = E.cpu().numpy(), V.cpu().numpy(), info.cpu().numpy()
            elif mode == "matel":
                r = eng.sweep(params, k, want_vectors=True, matrix_elements=True)
                E, V, info = r.E, r.V, r.info
                Mo = orc.matrix_elements(params[idx], r.V[idx].transpose(0, 2, 1))
                assert np.max(np.abs(r.matel[idx] - Mo)) < 1e-10 * max(1.0, np.abs(Mo).max())
            else:
                r = eng.sweep(params, k, want_vectors=True)
                E, V, info = r.E, r.V, r.info
        assert not info.any(), mode
        assert eig_rel_err(E[idx], ref["E"][:, :k]) < E_RTOL, mode
        assert np.abs(V[650].imag).max() > 1e-3                                          # genuinely complex eigenvectors
        for a, p in enumerate(idx):
            assert subspace_overlap(V[p].T, ref["V"][a][:, :k], ref["E"][a][:k], E_next=ref["E"][a][k]) >= OVERLAP, (mode, p)


def test_pauli_coefficients_fixed_operator_on_device(engine_mod):
    """util.pauliCoefficients with ONE operator for all points (util.py:320-342): the 2 x 2 matrix elements come from
    scqed_dense_matrix_elements (own kernel), then the reduction kernel; against the reference's stored hx, hy, hz
    (complex64 accuracy) and the float64 oracle on the same eigenvectors."""
    from oracle.plan_oracle import pauli_coefficients
    import pycqed_b200.util as util
    from pycqed_b200.numerical_system import KetGrid
    g, plan = load_golden("c6_rfsquid")
    k = int(g["k"])
    A = g["pauli_operator"]
    with engine_mod.Engine(plan) as eng:
        res = eng.sweep(g["params"], k, want_vectors=True)
    N = len(g["params"])
    M = engine_mod.dense_matrix_elements(A, res.V[:, :3])
    Mref = np.einsum("pai,ij,pbj->pab", res.V[:, :3].conj(), A, res.V[:, :3])
    assert np.max(np.abs(M - Mref)) < 1e-12 * max(1.0, np.abs(A).max())
    pt, lv = np.meshgrid(np.arange(N), np.arange(k), indexing="xy")
    V = KetGrid(lambda: res.V, pt, lv, [[plan.hilbert_dim], [1]])                       # what getSweep returns (k, N)
    hx, hy, hz = util.pauliCoefficients(res.E.T.copy(), V, A)
    ox, oy, oz = pauli_coefficients(res.E[:, 0], res.E[:, 1], Mref[:, :2, :2])
    scale = np.max(np.abs(res.E[:, :2]))
    for a, b in ((hx, ox), (hy, oy), (hz, oz)):
        assert np.max(np.abs(a - b)) < 1e-11 * scale
    # gauge invariant part against the reference's own numbers (it rounds to complex64)
    assert np.max(np.abs(np.hypot(hx, hy) - np.hypot(g["pauli_hx"], g["pauli_hy"]))) < 2e-5 * scale
    assert np.max(np.abs(hz - g["pauli_hz"])) < 2e-5 * scale
    # the same call with a plain object array of Ket-like vectors (what a reference user would hold)
    Vobj = np.asarray(V)
    h2 = util.pauliCoefficients(res.E.T.copy(), Vobj, A)
    assert all(np.array_equal(a, b) for a, b in zip((hx, hy, hz), h2))
