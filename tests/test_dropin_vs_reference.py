"""The drop-in ``pycqed_b200.NumericalSystem`` against the reference ``pyscqed.NumericalSystem``
driven through the same user-level calls (configureOperator / setParameterValues / addSweep /
addEvaluation / setDiagConfig / paramSweep / getSweep).

The build container has the reference but no GPU, the GPU box has a GPU but no reference, so the
engine is replaced HERE (test only) by a stand-in that answers the same ``Engine`` calls from the
CPU oracle.  What is under test is everything around the kernels: lowering, the evaluable
dispatch, the result container and the ``getSweepResult`` layout (parameters.py:1002-1166).
The kernels themselves are tested against the same golden data in tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

from tests.conftest import needs_reference

pytestmark = [needs_reference]


class _FakeTensor:
    def __init__(self, a):
        self._a = a

    def cpu(self):
        return self

    def numpy(self):
        return self._a

    def __getitem__(self, i):
        return _FakeTensor(self._a[i])


class OracleEngine:
    """Test double for pycqed_b200.engine.Engine (same call signatures), backed by the oracle."""

    def __init__(self, plan, device=None):
        from oracle.plan_oracle import PlanOracle
        self.plan = plan
        self.orc = PlanOracle(plan)
        self.n = plan.hilbert_dim

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def coefficients(self, params):
        from oracle.plan_oracle import eval_program
        c, s = eval_program(self.plan.program, np.atleast_2d(np.asarray(params, dtype=np.float64)))
        return _FakeTensor(c), _FakeTensor(s)

    def assemble_dense(self, params, tidyup=True):
        c, _ = self.coefficients(params)
        return _FakeTensor(np.stack([self.orc.hamiltonian(row, tidy=tidyup).toarray() for row in c.numpy()]))

    device = None

    def sweep(self, params, k, want_vectors=False, resonator=False, matrix_elements=False, **kw):
        from pycqed_b200.engine import SweepResult
        out = self.orc.sweep(params, k, get_vectors=want_vectors or matrix_elements, resonator=resonator)
        V = np.ascontiguousarray(out["V"].transpose(0, 2, 1)) if want_vectors else None
        scal = {nm: out["scalars"][:, i] for i, nm in enumerate(self.plan.program.scalar_names)}
        M = self.orc.matrix_elements(params, out["V"]) if matrix_elements else None
        return SweepResult(out["E"], V, scal, out["Erwa"], np.zeros(len(out["E"]), dtype=np.int32), M)

    def matrix_elements(self, params, V):
        V = V if isinstance(V, np.ndarray) else V.numpy()
        return _FakeTensor(self.orc.matrix_elements(params, np.asarray(V).transpose(0, 2, 1)))


class OraclePool:
    """Test double for pycqed_b200.engine.EnginePool: one oracle-backed engine."""

    def __init__(self, plan, devices=None):
        self.plan = plan
        self.engines = [OracleEngine(plan)]
        self.n = plan.hilbert_dim

    def close(self):
        pass

    def sweep(self, params, k, want_vectors=False, resonator=False, **kw):
        kw.pop("real_vectors", None)
        kw.pop("vectors_on_device", None)
        return self.engines[0].sweep(params, k, want_vectors=want_vectors, resonator=resonator, **kw)


@pytest.fixture()
def patched(monkeypatch):
    import pycqed_b200.numerical_system as ns
    monkeypatch.setattr(ns._engine, "Engine", OracleEngine)
    monkeypatch.setattr(ns._engine, "EnginePool", OraclePool)
    monkeypatch.setattr(ns._engine, "visible_devices", lambda: [0])
    return ns


def _pair(builder, patched, *args):
    """(reference system, drop-in system) configured identically on separate symbolic systems."""
    from tests.golden import make_golden as mg
    ref = getattr(mg, builder)(*args)
    twin = getattr(mg, builder)(*args)           # same circuit / parameterisations / values
    with mg.quiet():
        mine = patched.NumericalSystem(twin.SS, unit=twin.units)
        for node, data in twin.operator_data.items():
            mine.operator_data[node] = dict(data)
    return ref, mine, mg


def _run_both(ref, mine, mg, sweeps, k, get_vectors, evals):
    with mg.quiet():
        for h in (ref, mine):
            h.newSweep()
            for s in sweeps:
                h.addSweep(*s)
            for name, kw in evals:
                h.addEvaluation(name, **kw)
            h.setDiagConfig(eigvalues=k, get_vectors=get_vectors)
        a = ref.paramSweep(timesweep=False)
        b = mine.paramSweep(timesweep=False)
    return a, b


def test_1d_sweep_layout_and_values(patched):
    ref, mine, mg = _pair("build_transmon", patched, 12)
    a, b = _run_both(ref, mine, mg, [("phiZ", -0.4, 0.45, 7)], 6, False, [])
    xa, Ya, va = ref.getSweep(a, "phiZ", {})
    xb, Yb, vb = mine.getSweep(b, "phiZ", {})
    assert Ya.shape == Yb.shape == (6, 7)
    assert np.allclose(xa, xb) and va == vb == {}
    assert np.max(np.abs(Ya - Yb)) < 1e-10


def test_2d_sweep_static_and_mesh_layouts(patched):
    ref, mine, mg = _pair("build_averin", patched, 8)
    ev = [("Hamiltonian", {}), ("ChargingEnergy", {"node": 1}), ("JosephsonEnergy", {"edge": (0, 1, 2)})]
    a, b = _run_both(ref, mine, mg, [("Qc", -1.0, 1.0, 5), ("phie", -0.4, 0.4, 4)], 5, False, ev)
    # 1-D cut at the nearest static value (parameters.py:1081-1085)
    xa, Ya, va = ref.getSweep(a, "phie", {"Qc": 0.3})
    xb, Yb, vb = mine.getSweep(b, "phie", {"Qc": 0.3})
    assert Ya.shape == Yb.shape == (5, 4) and va == vb
    assert np.max(np.abs(Ya - Yb)) < 1e-10
    xa, Ya, va = ref.getSweep(a, "Qc", {"phie": -0.1})
    xb, Yb, vb = mine.getSweep(b, "Qc", {"phie": -0.1})
    assert Ya.shape == Yb.shape == (5, 5) and va == vb
    assert np.max(np.abs(Ya - Yb)) < 1e-10
    # mesh retrieval: (k, N_last, N_first) (parameters.py:1146-1157)
    xa, Za, _ = ref.getSweep(a, ["Qc", "phie"], {})
    xb, Zb, _ = mine.getSweep(b, ["Qc", "phie"], {})
    assert Za.shape == Zb.shape == (5, 4, 5)
    assert np.max(np.abs(Za - Zb)) < 1e-10
    assert all(np.allclose(p, q) for p, q in zip(xa, xb))
    # scalar evaluables
    for evn in ("ChargingEnergy", "JosephsonEnergy"):
        _, Sa, _ = ref.getSweep(a, ["Qc", "phie"], {}, evaluable=evn)
        _, Sb, _ = mine.getSweep(b, ["Qc", "phie"], {}, evaluable=evn)
        assert Sa.shape == Sb.shape == (4, 5)
        assert np.max(np.abs(Sa - Sb)) < 1e-12 * max(1.0, np.abs(Sa).max())


def test_vectors_and_resonator_layouts(patched):
    ref, mine, mg = _pair("build_fluxonium", patched, 30)
    ev = [("Hamiltonian", {}), ("Resonator", {"cpl_node": 1, "nmax": 12})]
    a, b = _run_both(ref, mine, mg, [("phiZ", -0.5, 0.5, 5)], 6, True, ev)
    xa, EVa, _ = ref.getSweep(a, "phiZ", {})
    xb, EVb, _ = mine.getSweep(b, "phiZ", {})
    assert EVa.shape == EVb.shape == (6, 2, 5)                    # EV[:, 0] = E, EV[:, 1] = V
    assert np.max(np.abs(EVa[:, 0].astype(float) - EVb[:, 0].astype(float))) < 1e-10
    for i in range(6):
        for p in range(5):
            va = EVa[i, 1, p].data.to_array()[:, 0]
            vb = EVb[i, 1, p].full()[:, 0]
            assert EVb[i, 1, p].dims == EVa[i, 1, p].dims
            assert abs(abs(np.vdot(va, vb)) - 1.0) < 1e-8
    _, Ra, _ = ref.getSweep(a, "phiZ", {}, evaluable="Resonator")
    _, Rb, _ = mine.getSweep(b, "phiZ", {}, evaluable="Resonator")
    assert Ra.shape == Rb.shape == (6, 12 + 1 + 6, 5)
    assert np.max(np.abs(Ra - Rb)) < 1e-9 * np.abs(Ra).max()
    # per-point records keep the reference's format
    rec = b.record(0)
    assert set(rec) == {"getHamiltonian", "getResonatorResponse"}
    assert isinstance(rec["getHamiltonian"][0], tuple) and len(rec["getHamiltonian"][1]) == 6


def test_current_voltage_evaluables(patched):
    """examples/local-basis-example.py:66-85,176-190: Current / Voltage in a sweep, getSweep layout
    (n_elements, N); diagonal elements are phase independent and must equal the reference's."""
    ref, mine, mg = _pair("build_rfsquid", patched, 30)
    jj = ref.getCircuitGraph().getComponentEdge("I")
    el = [(0, 0), (1, 1), (0, 1)]
    ev = [("Hamiltonian", {}), ("Current", {"edge": jj, "elements": el}), ("Voltage", {"node": 1, "elements": el})]
    a, b = _run_both(ref, mine, mg, [("phiZ", 0.49, 0.51, 5)], 4, True, ev)
    for evn in ("Current", "Voltage"):
        _, Ia, _ = ref.getSweep(a, "phiZ", {}, evaluable=evn)
        _, Ib, _ = mine.getSweep(b, "phiZ", {}, evaluable=evn)
        assert Ia.shape == Ib.shape == (3, 5)
        scale = max(1.0, np.abs(Ia).max())
        assert np.max(np.abs(Ia[:2] - Ib[:2])) < 1e-9 * scale
        assert np.max(np.abs(np.abs(Ia[2]) - np.abs(Ib[2]))) < 1e-7 * scale       # real eigenvectors: sign only
    rec = b.record(0)
    assert set(rec) == {"getHamiltonian", "getCurrentMatrixElement", "getVoltageMatrixElement"}
    assert rec["getCurrentMatrixElement"].shape == (3,)
    # error behaviour (numerical_system.py:598-601,660-661,671-674)
    for h in (ref, mine):
        with mg.quiet():
            h.newSweep()
            h.addSweep("phiZ", 0.49, 0.51, 2)
            h.addEvaluation("Hamiltonian")
            h.addEvaluation("Current", edge=(7, 8, 0), elements=el)
            h.setDiagConfig(eigvalues=3, get_vectors=True)
        with pytest.raises(Exception, match="not in the circuit"):
            with mg.quiet():
                h.paramSweep(timesweep=False)
        with mg.quiet():
            h.newSweep()
            h.addSweep("phiZ", 0.49, 0.51, 2)
            h.addEvaluation("Hamiltonian")
            h.addEvaluation("Voltage", node=1)
            h.setDiagConfig(eigvalues=3, get_vectors=True)
        with pytest.raises(Exception, match="No elements specified"):
            with mg.quiet():
                h.paramSweep(timesweep=False)


def test_current_voltage_operators_single_point(patched):
    ref, mine, mg = _pair("build_csfq_3jj", patched, 3)
    e12 = ref.getCircuitGraph().getComponentEdge("I2")
    with mg.quiet():
        for edge in (e12, ref.getCircuitGraph().getComponentEdge("I1")):
            A = ref.getCurrentOperator(edge).data.to_array()
            B = mine.getCurrentOperator(edge).full()
            assert np.max(np.abs(A - B)) < 1e-12 * max(1.0, np.abs(A).max())
        A = ref.getVoltageOperator(2).data.to_array()
        B = mine.getVoltageOperator(2).full()
        assert np.max(np.abs(A - B)) < 1e-12 * max(1.0, np.abs(A).max())
        ref.setDiagConfig(eigvalues=4, get_vectors=True)
        E, V = ref.diagonalize(ref.getHamiltonian())
        el = [(0, 0), (0, 1), (2, 3)]
        ma = ref.getCurrentMatrixElement(E, V, edge=e12, elements=el)
        mb = mine.getCurrentMatrixElement(E, V, edge=e12, elements=el)
        assert np.max(np.abs(ma - mb)) < 1e-10 * max(1.0, np.abs(ma).max())
        va = ref.getVoltageMatrixElement(E, V, node=1, elements=el)
        vb = mine.getVoltageMatrixElement(E, V, node=1, elements=el)
        assert np.max(np.abs(va - vb)) < 1e-10 * max(1.0, np.abs(va).max())
    ref2, mine2, mg = _pair("build_rfsquid", patched, 20)
    le = ref2.getCircuitGraph().getComponentEdge("L")
    with mg.quiet():
        A = ref2.getCurrentOperator(le).data.to_array()
        B = mine2.getCurrentOperator(le).full()
    assert np.max(np.abs(A - B)) < 1e-11 * max(1.0, np.abs(A).max())
    for h in (ref2, mine2):
        with pytest.raises(Exception, match="No edge specified"):
            h.getCurrentOperator()
        with pytest.raises(Exception, match="No node specified"):
            h.getVoltageOperator()


def test_single_point_api(patched):
    ref, mine, mg = _pair("build_csfq_3jj", patched, 3)
    with mg.quiet():
        Ha = ref.getHamiltonian().data.to_array()
        Hb = mine.getHamiltonian()
    assert Hb.shape == Ha.shape and Hb.dims == [[7, 7], [7, 7]]
    assert np.max(np.abs(Ha - Hb.full())) < 4e-12 * np.abs(Ha).max()
    assert mine.getHilbertSpaceSize() == ref.getHilbertSpaceSize() == 49
    with mg.quiet():
        assert abs(mine.getChargingEnergies(1) - ref.getChargingEnergies(1)) < 1e-12
        assert abs(mine.getJosephsonEnergies((1, 0, 0)) - ref.getJosephsonEnergies((1, 0, 0))) < 1e-12 \
            if (1, 0, 0) in ref.SS.edges else True
        ea, eb = ref.getJosephsonEnergies(), mine.getJosephsonEnergies()
    assert set(ea) == set(eb)
    assert all(abs(ea[k] - eb[k]) < 1e-12 for k in ea)
    # circ_operators accessor (examples/csfq-example.py:55)
    qa = ref.circ_operators[1]["charge"].data.to_array()
    assert np.max(np.abs(qa - mine.circ_operators[1]["charge"].full())) < 1e-12


def test_single_point_conveniences(patched):
    ref, mine, mg = _pair("build_fluxonium", patched, 18)
    with mg.quiet():
        ref.setDiagConfig(eigvalues=5, get_vectors=True)
        E, V = ref.diagonalize(ref.getHamiltonian())
        Ra = ref.getResonatorResponse(np.asarray(E), V, nmax=9, cpl_node=1)
        Rb = mine.getResonatorResponse(np.asarray(E), V, nmax=9, cpl_node=1)
        assert mine.getResonatorResponse(np.asarray(E), V) is None
    assert Ra.shape == Rb.shape == (9 + 1 + 5, 5)
    assert np.max(np.abs(Ra - Rb)) < 1e-9 * np.abs(Ra).max()
    with mg.quiet():
        la = [x.data.to_array() for x in ref.getOperatorList(1)]
        lb = mine.getOperatorList(1)
    for a, b in zip(la, lb):
        assert np.max(np.abs(a - b)) < 1e-11 * max(1.0, np.abs(a).max())
    with mg.quiet():
        sa, sb = ref.sparsity(ref.getHamiltonian()), mine.sparsity(mine.getHamiltonian())
    assert abs(sa - sb) < 1e-12
    mine.prepareOperators()


def test_error_behaviour_matches_reference(patched):
    ref, mine, mg = _pair("build_transmon", patched, 5)
    for h in (ref, mine):
        with pytest.raises(Exception, match="not valid"):
            h.addEvaluation("Nonsense")
        with pytest.raises(Exception, match="not a valid circuit node"):
            h.configureOperator(99, 5, "charge")
        with pytest.raises(Exception, match="not a qutip Qobj"):
            h.diagonalize(np.eye(3))
        with pytest.raises(ValueError):
            h.addSweep("nope", 0, 1, 3)
    # Resonator without eigenvectors: the reference raises "need eigenvectors for 'depends'"
    ref2, mine2, mg = _pair("build_fluxonium", patched, 12)
    for h in (ref2, mine2):
        with mg.quiet():
            h.newSweep()
            h.addSweep("phiZ", 0.1, 0.2, 2)
            h.addEvaluation("Hamiltonian")
            h.addEvaluation("Resonator", cpl_node=1)
            h.setDiagConfig(eigvalues=4, get_vectors=False)
        with pytest.raises(Exception, match="need eigenvectors"):
            with mg.quiet():
                h.paramSweep(timesweep=False)


def test_swept_oscillator_impedance_matches_reference(patched):
    """Sweeping L (or C) of an oscillator-basis node changes its impedance: the reference regenerates the
    node operators at every point (numerical_system.py:378-393,435-437,1064-1109); the plan keeps
    constant unit operators in the eigenbasis of a + a^dagger instead.  Eigenvalues, eigenvectors
    (rotated back), resonator strips and matrix elements must agree."""
    ref, mine, mg = _pair("build_fluxonium", patched, 24)
    mine.regen_mode = "reference"          # the reference's loop leaves Q / P at the pre-sweep impedance
    L0 = ref.getParameterValue("L")
    ev = [("Hamiltonian", {}), ("Resonator", {"cpl_node": 1, "nmax": 8}),
          ("Voltage", {"node": 1, "elements": [(0, 0), (1, 1), (0, 1)]})]
    a, b = _run_both(ref, mine, mg, [("L", 0.8 * L0, 1.3 * L0, 4), ("phiZ", 0.1, 0.5, 3)], 5, True, ev)
    xa, EVa, _ = ref.getSweep(a, ["L", "phiZ"], {})
    xb, EVb, _ = mine.getSweep(b, ["L", "phiZ"], {})
    assert EVa.shape == EVb.shape
    Ea, Eb = EVa[:, 0].astype(float), EVb[:, 0].astype(float)
    assert np.max(np.abs(Ea - Eb)) < 1e-9 * np.abs(Ea).max()
    for i in range(5):
        for idx in np.ndindex(*EVa.shape[2:]):
            va = EVa[(i, 1) + idx].data.to_array()[:, 0]
            vb = EVb[(i, 1) + idx].full()[:, 0]
            assert abs(abs(np.vdot(va, vb)) - 1.0) < 1e-8
    _, Ra, _ = ref.getSweep(a, ["L", "phiZ"], {}, evaluable="Resonator")
    _, Rb, _ = mine.getSweep(b, ["L", "phiZ"], {}, evaluable="Resonator")
    assert Ra.shape == Rb.shape and np.max(np.abs(Ra - Rb)) < 1e-9 * np.abs(Ra).max()
    _, Va, _ = ref.getSweep(a, ["L", "phiZ"], {}, evaluable="Voltage")
    _, Vb, _ = mine.getSweep(b, ["L", "phiZ"], {}, evaluable="Voltage")
    assert Va.shape == Vb.shape
    assert np.max(np.abs(Va[:2] - Vb[:2])) < 1e-9 * max(1.0, np.abs(Va).max())
    assert np.max(np.abs(np.abs(Va[2]) - np.abs(Vb[2]))) < 1e-7 * max(1.0, np.abs(Va).max())


def test_swept_impedance_consistent_mode_matches_single_point_path(patched):
    """regen_mode='consistent': at every sweep point H equals the reference's own single-point path
    (setParameterValues + getHamiltonian + diagonalize), which regenerates ALL operators of the node."""
    ref, mine, mg = _pair("build_fluxonium", patched, 20)
    mine.regen_mode = "consistent"
    C0 = ref.getParameterValue("C")
    with mg.quiet():
        mine.newSweep()
        mine.addSweep("C", 0.7 * C0, 1.2 * C0, 4)
        mine.setDiagConfig(eigvalues=5, get_vectors=True)
        data = mine.paramSweep(timesweep=False)
    x, EV, _ = mine.getSweep(data, "C", {})
    ref.setDiagConfig(eigvalues=5, get_vectors=True)
    for i, c in enumerate(x):
        with mg.quiet():
            ref.setParameterValue("C", float(c))
            E, V = ref.diagonalize(ref.getHamiltonian())
        assert np.max(np.abs(np.asarray(E) - EV[:, 0, i].astype(float))) < 1e-9 * np.abs(E).max()
        for j in range(5):
            assert abs(abs(np.vdot(V[j].data.to_array()[:, 0], EV[j, 1, i].full()[:, 0])) - 1.0) < 1e-8


def test_operator_factors_match_reference():
    """pycqed_b200.operators against numerical_system.py:1064-1210 (up to the diagonal-unitary gauge of
    ``eigenstates()`` signs, SURVEY.md Appendix D, which leaves |entries| and spectra unchanged)."""
    from tests.golden import make_golden as mg
    from pycqed_b200 import operators as ops
    h = mg.build_csfq_3jj(4)
    with mg.quiet():
        ref = [x.data.to_array() for x in h.getOperatorList(1)]
    mine = ops.charge_basis(4)
    for a, b in zip(ref, mine):
        assert np.max(np.abs(a - b)) < 1e-12
    f = mg.build_fluxonium(25)
    with mg.quiet():
        ref = [x.data.to_array() for x in f.getOperatorList(1)]
        z = f.getParameterValue("Zosc1") * f.units.getPrefactor("Impe")
    mine = ops.oscillator_basis(25, z, f.units.getPrefactor("ChgOsc"), f.units.getPrefactor("FlxOsc"),
                                f.SS.cooper_disp[1], f.SS.fluxon_disp[1])
    for a, b in zip(ref, mine):
        assert np.max(np.abs(a - b)) < 1e-11 * max(1.0, np.abs(a).max())
    with mg.quiet():
        h.operator_data[1]["basis"] = "flux"
        ref = [x.data.to_array() for x in h.getOperatorList(1)]
    mine = ops.flux_basis(4)
    for a, b in zip(ref, mine):
        assert np.max(np.abs(np.abs(a) - np.abs(b))) < 1e-12


def test_util_strip_helpers_match_reference():
    """pycqed_b200.util against pyscqed.util on a golden Resonator result (util.py:178-240, 286-298, 393-402)."""
    from tests.golden import make_golden as mg
    from tests.conftest import load_golden
    import pycqed_b200.util as mine
    ps = mg.import_reference()
    g, _ = load_golden("c2_fluxonium_res")
    for p in (0, 17, 40):
        Erwa = g["eval_getResonatorResponse"][p].T          # [level, strip] as getSweep(...)[:, :, p]
        assert np.array_equal(mine.getResonatorShift(Erwa), ps.util.getResonatorShift(Erwa))
        assert np.array_equal(mine.getCircuitLambShift(Erwa), ps.util.getCircuitLambShift(Erwa))
        na, sa = ps.util.getACStarkShift(Erwa)
        nb, sb = mine.getACStarkShift(Erwa)
        assert np.array_equal(na, nb) and np.array_equal(sa, sb)
    m = [np.linspace(0, 1, 5) + i for i in range(4)]
    ra = ps.util.createSubspaceOperators(*m)
    rb = mine.createSubspaceOperators(*m)
    assert rb.shape == (5, 2, 2)
    for a, b in zip(ra, rb):
        assert np.max(np.abs(np.asarray(a) - b)) < 1e-6      # the reference stores complex64
    sweep = np.empty((3, 2, 4), dtype=object)
    sweep[:, 0, :] = 1.5
    sweep[:, 1, :] = "ket"
    Ea, Va = ps.util.getEigenValuesAndVectors(sweep)
    Eb, Vb = mine.getEigenValuesAndVectors(sweep)
    assert np.array_equal(Ea, Eb) and Eb.dtype == np.float64 and Vb.shape == Va.shape


def test_all_nodes_scalar_evaluables_and_resonator_without_node(patched):
    """``addEvaluation('ChargingEnergy')`` without ``node=`` gives a dict per point (numerical_system.py:696-724);
    ``Resonator`` without ``cpl_node`` gives None records (:737-741).  Same records / getSweep output as the reference."""
    ref, mine, mg = _pair("build_csfq_3jj", patched, 3)
    ev = [("Hamiltonian", {}), ("ChargingEnergy", {}), ("JosephsonEnergy", {}), ("FluxEnergy", {"node": 2}),
          ("Resonator", {})]
    a, b = _run_both(ref, mine, mg, [("phiZ", 0.4, 0.6, 3)], 4, True, ev)
    recs_a = [mg.import_reference().util.pickleRead(f) for f in a]
    for i, ra in enumerate(recs_a):
        rb = b.record(i)
        assert set(ra) == set(rb)
        assert rb["getResonatorResponse"] is None and ra["getResonatorResponse"] is None
        for meth in ("getChargingEnergies", "getJosephsonEnergies"):
            assert set(ra[meth]) == set(rb[meth])
            for kk in ra[meth]:
                assert abs(ra[meth][kk] - rb[meth][kk]) <= 1e-13 * max(1.0, abs(ra[meth][kk]))
        assert abs(ra["getFluxEnergies"] - rb["getFluxEnergies"]) <= 1e-13
    for evn in ("ChargingEnergy", "JosephsonEnergy", "Resonator"):
        _, Sa, _ = ref.getSweep(a, "phiZ", {}, evaluable=evn)
        _, Sb, _ = mine.getSweep(b, "phiZ", {}, evaluable=evn)
        assert Sa.shape == Sb.shape == (3,)
        for x, y in zip(Sa, Sb):
            if x is None:
                assert y is None
            else:
                assert set(x) == set(y) and all(abs(x[q] - y[q]) <= 1e-13 * max(1.0, abs(x[q])) for q in x)


def test_plan_cache_and_lazy_results(patched):
    """Row f2: the second sweep of the same circuit re-uses the lowered plan and the uploaded engine (no sympy);
    getSweep on 10^4 x 10 eigenvectors creates no per-ket objects."""
    import time
    ref, mine, mg = _pair("build_fluxonium", patched, 16)
    calls = []
    orig = patched.lower_system

    def counting(*a, **kw):
        calls.append(1)
        return orig(*a, **kw)
    patched.lower_system = counting
    try:
        with mg.quiet():
            for N in (5, 7):
                mine.newSweep()
                mine.addSweep("phiZ", -0.5, 0.5, N)
                mine.addEvaluation("Hamiltonian")
                mine.addEvaluation("Resonator", cpl_node=1, nmax=6)
                mine.setDiagConfig(eigvalues=4, get_vectors=True)
                data = mine.paramSweep(timesweep=False)
                assert len(data) == N
            assert len(calls) == 1                      # the 7-point sweep re-used the plan of the 5-point one
            assert data.timings["init"] < 0.05
            H1 = mine.getHamiltonian()
            H2 = mine.getHamiltonian()
            assert len(calls) == 2 and np.array_equal(H1.full(), H2.full())
            mine.setParameterValue("phiZ", 0.25)        # a different configuration: lowered again
            mine.getHamiltonian()
            assert len(calls) == 3
    finally:
        patched.lower_system = orig
    x, EV, _ = mine.getSweep(data, "phiZ", {})
    assert EV.shape == (4, 2, 7)
    E, V = EV[:, 0], EV[:, 1]
    assert E.shape == (4, 7) and E.dtype == np.float64 and V.shape == (4, 7)
    assert abs(V[2, 3].norm() - 1.0) < 1e-12 and V[2, 3].dims == [[16], [1]]
    full = np.asarray(EV)
    assert full.shape == (4, 2, 7) and full[1, 0, 2] == E[1, 2] and isinstance(full[1, 1, 2], patched.Ket)
    # timing of the retrieval on a bench-sized synthetic result
    from pycqed_b200.engine import SweepResult
    n_pts, k, n = 10000, 10, 101
    res = SweepResult(np.zeros((n_pts, k)), np.zeros((n_pts, k, n)), {}, None, np.zeros(n_pts, dtype=np.int32))
    big = patched.SweepData([{"name": "phiZ", "start": -1.0, "end": 1.0, "N": n_pts}], ["getHamiltonian"], res, {}, [n], {})
    t0 = time.perf_counter()
    x, EV, _ = mine.getSweep(big, "phiZ", {})
    E, V = EV[:, 0], EV[:, 1]
    dt = time.perf_counter() - t0
    assert E.shape == (k, n_pts) and V.shape == (k, n_pts) and dt < 0.02, dt
