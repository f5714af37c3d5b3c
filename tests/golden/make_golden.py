#!/usr/bin/env python
"""Mint the golden fixtures by running the UNMODIFIED reference (``/root/reference/pyscqed``).

The reference has no tests / known answers for its sweep path, and it cannot
travel to the GPU box, so its outputs are frozen here:

  tests/golden/<cfg>.plan.json   the lowered SweepPlan (factor specs, term table, coefficient
                                 program) -- input of the GPU path, the oracle and bench.py
  tests/golden/<cfg>.npz         reference outputs on a reduced grid: params, ``E`` from
                                 ``paramSweep`` + ``util.diagDenseH`` (or eigsh tol=0 for n>=9261),
                                 a few Hamiltonians (CSR triplets), ``_postsub`` floats, eigenvector
                                 samples, resonator ``Erwa``, scalar evaluables

QuTiP is not installed in this image; the reference runs over ``oracle/shims`` (a scipy.sparse
restatement of the QuTiP calls it makes).  Circuits and parameter values are those of the
reference's example scripts (file:line cited per config).

Usage:  python tests/golden/make_golden.py [cfg ...]
"""
import contextlib
import io
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PYSCQED_REFERENCE", "/root/reference")


def import_reference():
    """Import the reference package over the compatibility shims (test infrastructure)."""
    warnings.filterwarnings("ignore")
    for p in (os.path.join(ROOT, "oracle", "shims"), REF, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    with contextlib.redirect_stdout(io.StringIO()):
        import pyscqed
    return pyscqed


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


# ------------------------------------------------------------------------------------------------
# the five BASELINE.json configs, built with the reference API
# ------------------------------------------------------------------------------------------------
def build_transmon(trunc=40):
    """C1: examples/transmon-example.py:20-37 plus a gate charge (examples/jpsq-example.py:35 style)."""
    ps = import_reference()
    with quiet():
        g = ps.CircuitGraph()
        g.addBranch(0, 1, "C")
        g.addBranch(0, 1, "I1")
        g.addBranch(0, 1, "I2")
        g.addFluxBias("I1", "Z")
        g.addChargeBias(1, "g", "Cg")
        h = ps.NumericalSystem(ps.SymbolicSystem(g))
        h.configureOperator(1, trunc, "charge")
        h.setParameterValues("C", 64.568, "I1", 16.1132e-3 / 2, "I2", 16.1132e-3 / 2,
                             "Cg", 0.0, "phiZ", 0.0, "Qg", 0.0)
    return h


def build_fluxonium(trunc=101):
    """C2: examples/fluxonium-example.py:25-73 (qubit A + capacitively coupled resonator)."""
    ps = import_reference()
    pc = ps.physical_constants
    EcA, ElA, EjA, frA, ZrA = 1.398e9, 0.523e9 * 0.5, 2.257e9, 6.696, 50.0
    C = 0.5 * pc.e ** 2 / (pc.h * EcA)
    L = 0.5 * pc.phi0 ** 2 / (pc.h * ElA * 4 * np.pi ** 2)
    I = 2 * np.pi * pc.h * EjA / pc.phi0
    with quiet():
        g = ps.CircuitGraph()
        g.addBranch(0, 1, "C")
        g.addBranch(0, 1, "L")
        g.addBranch(0, 1, "I")
        g.addFluxBias("I", "Z")
        g.coupleResonatorCapacitively(1, "Cc")
        h = ps.NumericalSystem(ps.SymbolicSystem(g))
        h.configureOperator(1, trunc, "oscillator")
        h.setParameterValues("C", C * 1e15 - 0.3, "I", I * 1e6, "L", L * 1e12, "Cc", 0.3,
                             "f1r", frA, "Z1r", ZrA, "phiZ", 0.5)
    return h


_FAB = dict(Ca=60.0, Jc=3.0, Aj=0.3 ** 2)


def build_csfq_3jj(trunc=10):
    """C3a: grounded 3JJ CSFQ, 2 nodes (examples/csfq-example.py:156-205)."""
    ps = import_reference()
    Ca, Jc, Aj = _FAB["Ca"], _FAB["Jc"], _FAB["Aj"]
    with quiet():
        g = ps.CircuitGraph()
        g.addBranch(0, 1, "I1")
        g.addBranch(0, 1, "C1")
        g.addBranch(1, 2, "I2")
        g.addBranch(1, 2, "C2")
        g.addBranch(1, 2, "Csh")
        g.addBranch(0, 2, "I3")
        g.addBranch(0, 2, "C3")
        g.addFluxBias("I2", "Z")
        h = ps.NumericalSystem(ps.SymbolicSystem(g))
        h.configureOperator(1, trunc, "charge")
        h.configureOperator(2, trunc, "charge")
        h.setParameterValues("C1", Ca * Aj, "C2", Ca * Aj, "C3", Ca * Aj, "I1", Jc * Aj, "I2", Jc * Aj,
                             "I3", Jc * Aj, "Csh", 45, "phiZ", 0.5)
    return h


def build_csfq_4jj(trunc=10):
    """C3b: tunable 4JJ CSFQ, 2 nodes, junction asymmetry d (examples/csfq-example.py:726-847)."""
    ps = import_reference()
    import sympy as sy
    Ca, Jc, Aj = _FAB["Ca"], _FAB["Jc"], _FAB["Aj"]
    with quiet():
        g = ps.CircuitGraph()
        g.addBranch(0, 1, "C1")
        g.addBranch(0, 1, "I1")
        g.addBranch(1, 2, "C2a")
        g.addBranch(1, 2, "I2a")
        g.addBranch(1, 2, "C2b")
        g.addBranch(1, 2, "I2b")
        g.addBranch(1, 2, "Csh")
        g.addBranch(0, 2, "C3")
        g.addBranch(0, 2, "I3")
        g.addFluxBias("I3", "Z")
        g.addFluxBias("I2a", "Xa")
        g.addFluxBias("I2b", "Xb")
        h = ps.NumericalSystem(ps.SymbolicSystem(g))
        h.configureOperator(1, trunc, "charge")
        h.configureOperator(2, trunc, "charge")
        S = h.getSymbolicSystem()
        S.getSymbols("phiX", symbol_overrides=[sy.symbols(r"\phi_X")])
        sym = S.getSymbolList()
        S.addParameterisation("phiXa", 0.5 * sym["phiX"])
        S.addParameterisation("phiXb", -0.5 * sym["phiX"])
        S.getSymbols("d", "I2a'", "C2a'", "I2b'", "C2b'",
                     symbol_overrides=[sy.symbols(r"d"), sy.symbols(r"I'_{a2}"), sy.symbols(r"C'_{a2}"),
                                       sy.symbols(r"I'_{b2}"), sy.symbols(r"C'_{b2}")])
        sym = S.getSymbolList()
        d = sym["d"]
        S.addParameterisation("C2a", (1 + d) * sym["C2a'"])
        S.addParameterisation("I2a", (1 + d) * sym["I2a'"])
        S.addParameterisation("C2b", (1 - d) * sym["C2b'"])
        S.addParameterisation("I2b", (1 - d) * sym["I2b'"])
        h.setParameterValues("C1", Ca * Aj, "C2a'", 0.5 * Ca * Aj, "C2b'", 0.5 * Ca * Aj, "C3", Ca * Aj,
                             "I1", Jc * Aj, "I2a'", 0.5 * Jc * Aj, "I2b'", 0.5 * Jc * Aj, "I3", Jc * Aj,
                             "Csh", 45, "phiZ", 0.5, "phiX", 0.0, "d", 0.0)
    return h


def build_csfq_floating(trunc=7):
    """C3c: floating 3JJ CSFQ, 3 nodes (examples/csfq-example.py:432-479) with the alpha
    parameterisation C2 = alpha C2', I2 = alpha I2' (examples/csfq-example.py:326-337)."""
    ps = import_reference()
    import sympy as sy
    Ca, Jc, Aj = _FAB["Ca"], _FAB["Jc"], _FAB["Aj"]
    with quiet():
        g = ps.CircuitGraph()
        g.addBranch(0, 1, "Cgnd1")
        g.addBranch(0, 2, "Cgnd2")
        g.addBranch(0, 3, "Cgnd3")
        g.addBranch(1, 2, "C1")
        g.addBranch(1, 2, "I1")
        g.addBranch(2, 3, "C2")
        g.addBranch(2, 3, "I2")
        g.addBranch(2, 3, "Csh")
        g.addBranch(1, 3, "C3")
        g.addBranch(1, 3, "I3")
        g.addFluxBias("I1", "Z")
        h = ps.NumericalSystem(ps.SymbolicSystem(g))
        for n in (1, 2, 3):
            h.configureOperator(n, trunc, "charge")
        S = h.getSymbolicSystem()
        S.getSymbols("alpha", "C2'", "I2'",
                     symbol_overrides=[sy.symbols(r"\alpha"), sy.symbols(r"C'_{2}"), sy.symbols(r"I'_{2}")])
        sym = S.getSymbolList()
        S.addParameterisation("C2", sym["alpha"] * sym["C2'"])
        S.addParameterisation("I2", sym["alpha"] * sym["I2'"])
        h.setParameterValues("C1", Ca * Aj, "C2'", Ca * Aj, "C3", Ca * Aj, "I1", Jc * Aj, "I2'", Jc * Aj,
                             "I3", Jc * Aj, "Csh", 45, "Cgnd1", 2, "Cgnd2", 2, "Cgnd3", 2,
                             "phiZ", 0.5, "alpha", 0.44)
    return h


def build_averin(trunc=20):
    """C4: Averin coupler (examples/averin-coupler.py:40-44) with the fabrication parameterisation
    (:265-271) and the DC-SQUID critical current Ic = cos(pi phie) Jc wse w1 (:530-553)."""
    ps = import_reference()
    import sympy as sy
    with quiet():
        g = ps.CircuitGraph()
        g.addBranch(0, 1, "Cc")
        g.addBranch(0, 1, "C1")
        g.addBranch(0, 1, "Ic")
        g.addChargeBias(1, "c", "Cg1")
        h = ps.NumericalSystem(ps.SymbolicSystem(g))
        h.configureOperator(1, trunc, "charge")
        s = h.getSymbolicSystem()
        sym = s.getSymbols("Jc", "Ca", "wse", "w1")
        s.addParameterisation("Cc", sym["wse"] * sym["w1"] * sym["Ca"])
        sym2 = s.getSymbols("phie", symbol_overrides=[sy.symbols("\\Phi_e")])
        s.addParameterisation("Ic", sy.cos(sy.pi * sym2["phie"]) * s.getSymbol("Jc") * s.getSymbol("wse") * s.getSymbol("w1"))
        h.setParameterValues("wse", 0.1, "w1", 0.15, "Ca", 60, "Jc", 1.5, "Cg1", 1, "C1", 7.7,
                             "Qc", 0.5, "phie", 0.0)
    return h


def build_fluxonium_atom(trunc=31):
    """C5: inductively coupled fluxonium pair (examples/fluxonium-atom.py:26-81)."""
    ps = import_reference()
    pc = ps.physical_constants
    Ec, El, Ej = 3.4e9, 1.2e9 * 0.5, 9.4e9
    C = 0.5 * pc.e ** 2 / (pc.h * Ec) * 1e15
    L = 0.5 * pc.phi0 ** 2 / (pc.h * El * 4 * np.pi ** 2) * 1e12
    I = 2 * np.pi * pc.h * Ej / pc.phi0 * 1e6
    with quiet():
        g = ps.CircuitGraph()
        g.addBranch(0, 1, "L1")
        g.addBranch(0, 1, "I1")
        g.addBranch(0, 1, "C1")
        g.addBranch(0, 2, "L2")
        g.addBranch(0, 2, "I2")
        g.addBranch(0, 2, "C2")
        g.coupleBranchesInductively("L1", "L2", "M")
        g.addFluxBias("I1", "l")
        g.addFluxBias("I2", "r")
        h = ps.NumericalSystem(ps.SymbolicSystem(g))
        h.configureOperator(1, trunc, "oscillator")
        h.configureOperator(2, trunc, "oscillator")
        S = h.getSymbolicSystem()
        S.getSymbols("phi+", "phi-", "L")
        sym = S.getSymbolList()
        S.addParameterisation("phil", sym["phi+"] + sym["phi-"])
        S.addParameterisation("phir", sym["phi+"] - sym["phi-"])
        S.addParameterisation("M", sym["L"])
        S.addParameterisation("L1", 2 * sym["L"])
        S.addParameterisation("L2", 2 * sym["L"])
        h.setParameterValues("L", L, "C1", C, "I1", I, "C2", C, "I2", I, "phi+", 0.0, "phi-", 0.0)
    return h


def build_rfsquid(trunc=50):
    """C6a: RF-SQUID qubit of examples/local-basis-example.py:28-49 (oscillator basis)."""
    ps = import_reference()
    Ca, Jc, Aj = 60.0, 3.0, 0.2 * 1.2
    with quiet():
        g = ps.CircuitGraph()
        g.addBranch(0, 1, "C")
        g.addBranch(0, 1, "L")
        g.addBranch(0, 1, "I")
        g.addFluxBias("I", "Z")
        h = ps.NumericalSystem(ps.SymbolicSystem(g))
        h.configureOperator(1, trunc, "oscillator")
        h.setParameterValues("C", Ca * Aj, "I", Jc * Aj, "L", 570.0, "phiZ", 0.5)
    return h


def build_cpb(trunc=20):
    """C6b: Cooper-pair box of examples/local-basis-example.py:145-170 (charge basis, gate charge)."""
    ps = import_reference()
    Ca, Jc, Aj = 60.0, 1.0, 0.1 * 0.1
    with quiet():
        g = ps.CircuitGraph()
        g.addBranch(0, 1, "C")
        g.addBranch(0, 1, "I")
        g.addChargeBias(1, "Z", "Cg1")
        h = ps.NumericalSystem(ps.SymbolicSystem(g))
        h.configureOperator(1, trunc, "charge")
        h.setParameterValues("C", Ca * Aj, "I", Jc * Aj, "Cg1", 1.0, "QZ", 0.5)
    return h


# ------------------------------------------------------------------------------------------------
# running the reference and freezing its outputs
# ------------------------------------------------------------------------------------------------
def _csr_triplets(H):
    m = H.data.as_scipy().tocoo()
    return m.row.astype(np.int32), m.col.astype(np.int32), m.data.astype(np.complex128)


def reference_sweep(h, sweeps, k, get_vectors=False, evaluations=(), sparse_opts=None, timesweep=False):
    """``paramSweep`` of the reference exactly as a user would drive it
    (numerical_system.py:860-928); returns per-point records read back from its pickles."""
    ps = import_reference()
    with quiet():
        h.newSweep()
        for s in sweeps:
            h.addSweep(s["name"], s["start"], s["end"], s["N"])
        for name, kw in evaluations:
            h.addEvaluation(name, **kw)
        if sparse_opts is not None:
            h.setDiagConfig(eigvalues=k, get_vectors=get_vectors, sparse=True, sparsesolveropts=sparse_opts)
        else:
            h.setDiagConfig(eigvalues=k, get_vectors=get_vectors)
        t0 = time.time()
        files = h.paramSweep(timesweep=timesweep)
        dt = time.time() - t0
    recs = [ps.util.pickleRead(f) for f in files]
    for f in files:
        try:
            os.remove(f)
        except OSError:
            pass
    return recs, dt


def lower(h, sweeps, **kw):
    from pycqed_b200.lowering import lower_system
    with quiet():
        if sweeps:
            h.SS.ndSweep([h.SS.paramSweepSpec(s["name"], s["start"], s["end"], s["N"]) for s in sweeps])
    return lower_system(h.SS, h.units, h.operator_data, [s["name"] for s in sweeps], sweeps, **kw)


def freeze(name, h, sweeps, k=10, get_vectors=False, evaluations=(), sparse_opts=None, n_h_samples=3,
           vec_points=(), lower_kw=None, extra_plans=None, store_h=True, k_extra=0, vec_dtype=np.complex128):
    """Run the reference on ``sweeps`` and write <name>.npz + <name>.plan.json.

    ``k_extra``: the reference is asked for k + k_extra levels; ``E`` / ``V`` keep the lowest k and ``E_next`` the
    following ones, so that a test can tell whether level k-1 belongs to a degenerate cluster cut by the truncation.
    ``vec_dtype=np.complex64`` halves the stored eigenvectors of the big configurations (the overlap criterion
    1 - |<a|b>| is quadratic in the stored rounding error: ~1e-14)."""
    lower_kw = lower_kw or {}
    t0 = time.time()
    recs, dt = reference_sweep(h, sweeps, k + k_extra, get_vectors, evaluations, sparse_opts)
    out = {}
    multi = isinstance(recs[0], dict)
    ham = [r["getHamiltonian"] for r in recs] if multi else recs
    if get_vectors:
        E = np.array([np.asarray(r[0], dtype=np.float64) for r in ham])
        vec_points = list(vec_points) if len(vec_points) else [0]
        out["vec_points"] = np.asarray(vec_points, dtype=np.int64)
        V = np.array([np.stack([v.data.to_array()[:, 0] for v in ham[i][1]], axis=1) for i in vec_points])
        out["V"] = np.ascontiguousarray(V[:, :, :k]).astype(vec_dtype)
    else:
        E = np.array([np.asarray(r, dtype=np.float64) for r in ham])
    if k_extra:
        out["E_next"] = np.ascontiguousarray(E[:, k:])
        E = np.ascontiguousarray(E[:, :k])
    out["E"] = E
    if multi:
        for key in recs[0]:
            if key != "getHamiltonian":
                out["eval_" + key] = np.array([np.asarray(r[key], dtype=np.float64) for r in recs])
    plan = lower(h, sweeps, **lower_kw)
    grid = plan.sweep_grid()
    out["params"] = grid
    out["swept_names"] = np.array(plan.swept_names)
    out["k"] = np.int64(k)
    out["ref_seconds_per_point"] = np.float64(dt / len(recs))
    # Hamiltonians + _postsub floats at a few points (assembly / coefficient parity)
    if store_h:
        idx = sorted(set(np.linspace(0, len(recs) - 1, n_h_samples).astype(int).tolist()))
        out["h_points"] = np.asarray(idx, dtype=np.int64)
        with quiet():
            h.SS.ndSweep([h.SS.paramSweepSpec(s["name"], s["start"], s["end"], s["N"]) for s in sweeps])
            h._presub()
            for j, i in enumerate(idx):
                h._postsub({kk: v[i] for kk, v in h.SS.sweep_grid_c.items()})
                H = h.getHamiltonian()
                r, c, v = _csr_triplets(H)
                out["H%d_row" % j], out["H%d_col" % j], out["H%d_val" % j] = r, c, v
                out["postsub%d_Cinv" % j] = np.asarray(h.Cinvnp, dtype=np.float64)
                out["postsub%d_Linv" % j] = np.asarray(h.Linvnp, dtype=np.float64)
                out["postsub%d_J" % j] = np.asarray(h.Jvecnp, dtype=np.float64)
                out["postsub%d_Pexp" % j] = np.asarray(h.Pexp_pnp, dtype=np.complex128)
    out["n"] = np.int64(plan.hilbert_dim)      # (the reference's getHilbertSpaceSize ignores 'flux' nodes, :93-96)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    plan.meta["source"] = name
    plan.save(os.path.join(HERE, name + ".plan.json"))
    for pname, (hh, sw, kw) in (extra_plans or {}).items():
        p2 = lower(hh, sw, **(kw or {}))
        p2.meta["source"] = pname
        p2.save(os.path.join(HERE, pname + ".plan.json"))
    print("%-28s n=%-6d pts=%-5d ref %.1f ms/pt  (%.1fs total)" % (
        name, int(out["n"]), len(recs), 1e3 * dt / len(recs), time.time() - t0), flush=True)


def S(name, a, b, n):
    return {"name": name, "start": a, "end": b, "N": n}


def make_c1():
    freeze("c1_transmon", build_transmon(40), [S("phiZ", -0.5, 0.5, 40), S("Qg", -0.5, 0.5, 25)], k=10)
    freeze("c1_transmon_vec", build_transmon(40), [S("phiZ", -0.5, 0.5, 5), S("Qg", -0.45, 0.45, 4)], k=10,
           get_vectors=True, vec_points=[0, 7, 19])


def make_c2():
    freeze("c2_fluxonium", build_fluxonium(101), [S("phiZ", -1.0, 1.0, 201)], k=10,
           extra_plans={"c2_fluxonium_t200": (build_fluxonium(200), [S("phiZ", -1.0, 1.0, 10000)], None),
                        "c2_fluxonium_10k": (build_fluxonium(101), [S("phiZ", -1.0, 1.0, 10000)], None)})
    res = {"resonator": {"cpl_node": 1, "nmax": 100}}
    freeze("c2_fluxonium_res", build_fluxonium(101), [S("phiZ", -1.0, 1.0, 41)], k=20, get_vectors=True,
           evaluations=[("Hamiltonian", {}), ("Resonator", {"cpl_node": 1})], vec_points=[0, 10, 30, 31],
           lower_kw=res,
           extra_plans={"c2_fluxonium_res_10k": (build_fluxonium(101), [S("phiZ", -1.0, 1.0, 10000)], res)})
    freeze("c2_fluxonium_t200s", build_fluxonium(200), [S("phiZ", -1.0, 1.0, 9)], k=10)


def make_c2b():
    """C2 at truncation 200 with the Resonator evaluable: the post-op behind the HBM dense solver."""
    res = {"resonator": {"cpl_node": 1, "nmax": 30}}
    freeze("c2_fluxonium_res_t200s", build_fluxonium(200), [S("phiZ", -1.0, 1.0, 7)], k=10, get_vectors=True,
           evaluations=[("Hamiltonian", {}), ("Resonator", {"cpl_node": 1, "nmax": 30})], vec_points=[0, 3],
           lower_kw=res, store_h=False)


def make_c3():
    freeze("c3_csfq3jj_t10", build_csfq_3jj(10), [S("phiZ", 0.0, 1.0, 21)], k=10)
    freeze("c3_csfq4jj_t10", build_csfq_4jj(10), [S("d", 0.0, 0.1, 3), S("phiX", -1.0, 1.0, 11)], k=10,
           extra_plans={
               "c3_csfq4jj_t27": (build_csfq_4jj(27), [S("d", 0.0, 0.1, 16), S("phiX", -1.0, 1.0, 64)], None),
               "c3_csfq4jj_t50": (build_csfq_4jj(50), [S("d", 0.0, 0.1, 16), S("phiX", -1.0, 1.0, 64)], None)})
    freeze("c3_csfq4jj_t27s", build_csfq_4jj(27), [S("d", 0.0, 0.1, 2), S("phiX", -1.0, 1.0, 2)], k=10,
           store_h=False)
    # floating 3-node: exact doublets at phiZ = 0.5 -> vectors stored
    freeze("c3_csfqfloat_t5", build_csfq_floating(5), [S("alpha", 0.3, 1.0, 3), S("phiZ", 0.45, 0.55, 5)], k=10,
           get_vectors=True, vec_points=[2, 7, 9])
    freeze("c3_csfqfloat_t7", build_csfq_floating(7), [S("alpha", 0.3, 1.0, 2), S("phiZ", 0.45, 0.55, 3)], k=10,
           n_h_samples=2,
           extra_plans={"c3_csfqfloat_t7_full": (build_csfq_floating(7),
                                                 [S("alpha", 0.3, 1.0, 16), S("phiZ", 0.45, 0.55, 64)], None),
                        "c3_csfqfloat_t10_full": (build_csfq_floating(10),
                                                  [S("alpha", 0.3, 1.0, 16), S("phiZ", 0.45, 0.55, 64)], None)})
    # n = 9261: the dense reference takes ~1 min/pt; use the reference's sparse path with tol=0
    # (examples/csfq-example.py:535-553; agrees with dense to 1e-14)
    freeze("c3_csfqfloat_t10", build_csfq_floating(10), [S("alpha", 0.44, 0.9, 2), S("phiZ", 0.47, 0.5, 2)], k=10,
           sparse_opts={"sigma": -300, "mode": "normal", "maxiter": None, "tol": 0}, n_h_samples=1)


_SA_EXACT = {"sigma": None, "mode": "normal", "maxiter": None, "tol": 0, "which": "SA"}


def make_big():
    """The two largest BASELINE configs through the reference's SPARSE path (util.diagSparseH, util.py:76-125) with
    ``tol=0`` -- ARPACK to machine precision, the setting SURVEY.md section 8d names for n >= 9261 -- since the dense
    path is out of reach (n = 32 400: 16.8 GB per matrix).  Eleven levels are requested, ten are the fixture."""
    freeze("c3_csfq4jj_t50s", build_csfq_4jj(50), [S("d", 0.02, 0.1, 2), S("phiX", -0.35, 0.8, 2)], k=10, k_extra=1,
           get_vectors=True, vec_points=[0, 3], sparse_opts=_SA_EXACT, store_h=False, vec_dtype=np.complex64)
    freeze("c5_fluxatom_t180s", build_fluxonium_atom(180), [S("phi-", 0.13, 0.5, 2)], k=10, k_extra=1,
           get_vectors=True, vec_points=[1], sparse_opts=_SA_EXACT, store_h=False, vec_dtype=np.complex64)


def build_phaseslip(basis, trunc):
    """a4: a phase-slip nanowire ("V" branch) shunted by L and C with a gate charge -- the dual of the Cooper-pair
    box.  Exercises the Hp terms (numerical_system.py:558-591: exp(+-2 pi i q) pdisp / pdisp_adj) and, with
    basis='flux', the flux basis of numerical_system.py:1159-1210 (P diagonal, Q dense, S = shift)."""
    ps = import_reference()
    with quiet():
        g = ps.CircuitGraph()
        g.addBranch(0, 1, "C")
        g.addBranch(0, 1, "L")
        g.addBranch(0, 1, "V")
        g.addChargeBias(1, "g", "Cg")
        h = ps.NumericalSystem(ps.SymbolicSystem(g))
        h.configureOperator(1, trunc, basis)
        h.setParameterValues("C", 20.0, "L", 800.0, "V", 30.0, "Cg", 1.0, "Qg", 0.3)
    return h


def make_a4():
    """flux-basis node and phase-slip branch through the whole path."""
    sw = [S("Qg", -0.45, 0.5, 9)]
    freeze("c9_phaseslip_flux", build_phaseslip("flux", 10), sw, k=5, k_extra=1, get_vectors=True, vec_points=[0, 4, 7])
    freeze("c9_phaseslip_osc", build_phaseslip("oscillator", 30), sw, k=5, k_extra=1, get_vectors=True, vec_points=[1, 6])
    # two sweep axes, the second one through the phase-slip amplitude itself
    freeze("c9_phaseslip_charge", build_phaseslip("charge", 10), [S("V", 5.0, 40.0, 3), S("Qg", -0.4, 0.4, 5)], k=5)


def _dropin_twin(h):
    """The drop-in class configured like reference system ``h`` (same symbolic system object)."""
    from pycqed_b200.numerical_system import NumericalSystem
    with quiet():
        mine = NumericalSystem(h.SS, unit=h.units, device=0)
        for node, data in h.operator_data.items():
            mine.operator_data[node] = dict(data)
    return mine


def make_single():
    """The single-point API -- getHamiltonian (numerical_system.py:509), diagonalize (:809), getResonatorResponse
    (:736), get*Energies (:696-734) at the current parameter values -- i.e. plans WITHOUT a swept column, lowered with
    the requests the drop-in class makes (``_single_plan``: all scalar evaluables; ``getResonatorResponse``)."""
    for name, h, k, res in (("c8_single_fluxonium", build_fluxonium(101), 20, {"cpl_node": 1, "nmax": 100}),
                            ("c8_single_csfq3jj", build_csfq_3jj(6), 8, None)):
        out = {}
        with quiet():
            h.setDiagConfig(eigvalues=k, get_vectors=True)
            H = h.getHamiltonian()
            E, V = h.diagonalize(H)
            out["E"] = np.asarray(E, dtype=np.float64)[None]
            out["V"] = np.stack([v.data.to_array()[:, 0] for v in V], axis=1)[None]
            out["vec_points"] = np.array([0])
            r, c, v = _csr_triplets(H)
            out["H0_row"], out["H0_col"], out["H0_val"] = r, c, v
            out["h_points"] = np.array([0])
            if res is not None:
                out["eval_getResonatorResponse"] = np.asarray(h.getResonatorResponse(np.asarray(E), V, nmax=res["nmax"], cpl_node=res["cpl_node"]))[None]
            ce, fe, je = h.getChargingEnergies(), h.getFluxEnergies(), h.getJosephsonEnergies()
        mine = _dropin_twin(h)
        scal = mine._scalar_exprs()
        plan = lower(h, [], scalars=scal)
        names = list(plan.program.scalar_names)
        ref_scal = np.zeros(len(names))
        for i, nm in enumerate(names):
            kind, key = nm.split(":", 1)
            key = eval(key)
            ref_scal[i] = {"ChargingEnergy": ce, "FluxEnergy": fe, "JosephsonEnergy": je}.get(kind, {}).get(key, np.nan)
        out["scalar_names"] = np.array(names)
        out["scalar_values"] = ref_scal
        out["params"] = np.zeros((1, 0))
        out["swept_names"] = np.array([], dtype=str)
        out["k"] = np.int64(k)
        out["n"] = np.int64(h.getHilbertSpaceSize())
        out["ref_seconds_per_point"] = np.float64(0.0)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        plan.meta["source"] = name
        plan.save(os.path.join(HERE, name + ".plan.json"))
        if res is not None:
            p2 = lower(h, [], resonator=res)
            p2.meta["source"] = name + "_res"
            p2.save(os.path.join(HERE, name + "_res.plan.json"))
        print("%-28s n=%d (single point)" % (name, int(out["n"])), flush=True)


def make_c4():
    ev = [("Hamiltonian", {}), ("ChargingEnergy", {"node": 1}), ("JosephsonEnergy", {"edge": (0, 1, 2)})]
    h = build_averin(20)
    S_ = h.getSymbolicSystem()
    scal = {"ChargingEnergy_1": None, "JosephsonEnergy_0": None}
    freeze("c4_averin", h, [S("Qc", -1.0, 1.0, 11), S("phie", -0.5, 0.5, 21)], k=10, evaluations=ev,
           lower_kw={"scalars": averin_scalars(h)},
           extra_plans={"c4_averin_full": (build_averin(20), [S("Qc", -1.0, 1.0, 101), S("phie", -0.5, 0.5, 201)],
                                           {"scalars": averin_scalars(h)})})


def averin_scalars(h):
    """ChargingEnergy(node=1) = 1/2 Cinv_00 Ec and JosephsonEnergy(edge) = J_e Ej
    (numerical_system.py:696-724) as coefficient-program outputs."""
    import sympy as sy
    Cinv = h.SS.getInverseCapacitanceMatrix()
    J = h.SS.getJosephsonVector()
    e = h.SS.edges.index((0, 1, 2))
    return {"ChargingEnergy": sy.Rational(1, 2) * Cinv[0, 0] * h.units.getPrefactor("Ec"),
            "JosephsonEnergy": J[e] * h.units.getPrefactor("Ej")}


def make_c5():
    freeze("c5_fluxatom_t31", build_fluxonium_atom(31), [S("phi-", 0.0, 1.0, 8)], k=10, n_h_samples=2,
           extra_plans={"c5_fluxatom_t31_64": (build_fluxonium_atom(31), [S("phi-", 0.0, 1.0, 64)], None),
                        "c5_fluxatom_t180": (build_fluxonium_atom(180), [S("phi-", 0.0, 1.0, 64)], None)})


def _pauli_reference(name, h, sweeps, k, evaluations, matel, basis):
    """Freeze a sweep with Current / Voltage evaluables plus util.pauliCoefficients
    (util.py:300-391; examples/local-basis-example.py:98-111,176-203) computed by the reference."""
    ps = import_reference()
    freeze(name, h, sweeps, k=k, get_vectors=True, evaluations=evaluations, vec_points=[0, 3],
           lower_kw={"matrix_elements": matel})
    recs, _ = reference_sweep(h, sweeps, k, True, evaluations)
    E = np.array([np.asarray(r["getHamiltonian"][0]) for r in recs]).T           # [k, N]
    V = np.empty((k, len(recs)), dtype=object)
    for i, r in enumerate(recs):
        V[:, i] = r["getHamiltonian"][1]
    with quiet():
        if basis[0] == "operator":        # a fixed operator at the initial parameter values
            h.setParameterValues(*basis[2])
            Op = h.getCurrentOperator(basis[1])
            op_dense = Op.data.to_array()
            hx, hy, hz = ps.util.pauliCoefficients(E, V, Op)
            m = np.array([[[Op.matrix_element(V[a, i], V[b, i]) for b in (0, 1)] for a in (0, 1)] for i in range(len(recs))])
        else:                              # subspace operators from the Voltage matrix elements
            Vii = np.array([r["getVoltageMatrixElement"] for r in recs]).T
            ops_ = ps.util.createSubspaceOperators(Vii[0], Vii[1], Vii[2], Vii[3])
            hx, hy, hz = ps.util.pauliCoefficients(E, V, ops_)
            m = np.array([np.asarray(o) for o in ops_])
            op_dense = np.zeros((0, 0))
    path = os.path.join(HERE, name + ".npz")
    g = dict(np.load(path))
    g.update(pauli_hx=hx, pauli_hy=hy, pauli_hz=hz, pauli_op=m.astype(np.complex128), pauli_operator=op_dense)
    np.savez_compressed(path, **g)


def make_c6():
    """Eigenvector consumers: Current / Voltage evaluables and pauliCoefficients."""
    h = build_rfsquid(50)
    jj = h.getCircuitGraph().getComponentEdge("I")
    le = h.getCircuitGraph().getComponentEdge("L")
    sw = [S("phiZ", 0.49, 0.51, 11)]
    el = [(0, 0), (0, 1), (1, 1), (2, 0)]
    _pauli_reference("c6_rfsquid", h, sw, 5,
                     [("Hamiltonian", {}), ("Current", {"edge": jj, "elements": el}), ("Voltage", {"node": 1, "elements": [(0, 1), (1, 2)]})],
                     [{"kind": "Current", "edge": list(jj), "elements": el}, {"kind": "Voltage", "node": 1, "elements": [(0, 1), (1, 2)]}],
                     ("operator", le, ("phiZ", 0.5)))
    h = build_rfsquid(50)
    freeze("c6_rfsquid_L", h, sw, k=5, get_vectors=True,
           evaluations=[("Hamiltonian", {}), ("Current", {"edge": le, "elements": el})], vec_points=[0, 3],
           lower_kw={"matrix_elements": [{"kind": "Current", "edge": list(le), "elements": el}]})
    h = build_cpb(20)
    el4 = [(0, 0), (0, 1), (1, 0), (1, 1)]
    _pauli_reference("c6_cpb", h, [S("QZ", 0.4, 0.6, 11)], 5,
                     [("Hamiltonian", {}), ("Voltage", {"node": 1, "elements": el4})],
                     [{"kind": "Voltage", "node": 1, "elements": el4}], ("subspace",))
    # multi-node: a junction between two non-ground nodes (two-factor current operator)
    h = build_csfq_3jj(5)
    e12 = h.getCircuitGraph().getComponentEdge("I2")
    freeze("c6_csfq3jj_t5", h, [S("phiZ", 0.45, 0.55, 5)], k=6, get_vectors=True,
           evaluations=[("Hamiltonian", {}), ("Current", {"edge": e12, "elements": el}), ("Voltage", {"node": 2, "elements": el})],
           vec_points=[0, 2],
           lower_kw={"matrix_elements": [{"kind": "Current", "edge": list(e12), "elements": el},
                                         {"kind": "Voltage", "node": 2, "elements": el}]})


def make_c7():
    """Swept oscillator impedance (numerical_system.py:378-393,435-437,1064-1109).

    c7_fluxonium_LphiZ   the reference's own sweep loop over L x phiZ (regen_mode='reference': its loop keeps Q / P
                         at the pre-sweep impedance), with eigenvectors and the Resonator evaluable
    c7_fluxonium_C       regen_mode='consistent': reference values from its single-point path
                         (setParameterValue + getHamiltonian + diagonalize at every grid point)"""
    h = build_fluxonium(40)
    L0 = h.getParameterValue("L")
    res = {"resonator": {"cpl_node": 1, "nmax": 20}, "regen_mode": "reference"}
    freeze("c7_fluxonium_LphiZ", h, [S("L", 0.8 * L0, 1.3 * L0, 5), S("phiZ", 0.0, 0.5, 4)], k=8, get_vectors=True,
           evaluations=[("Hamiltonian", {}), ("Resonator", {"cpl_node": 1, "nmax": 20})], vec_points=[0, 7, 19],
           lower_kw=res, store_h=False)
    # the pre-sweep impedance enters the plan in this mode, and the reference's sweep leaves the system at
    # the last grid point: lower again from a fresh system
    plan = lower(build_fluxonium(40), [S("L", 0.8 * L0, 1.3 * L0, 5), S("phiZ", 0.0, 0.5, 4)], **res)
    plan.meta["source"] = "c7_fluxonium_LphiZ"
    plan.save(os.path.join(HERE, "c7_fluxonium_LphiZ.plan.json"))
    h = build_fluxonium(40)
    C0 = h.getParameterValue("C")
    sw = [S("C", 0.7 * C0, 1.2 * C0, 9)]
    plan = lower(h, sw)
    grid = plan.sweep_grid()
    k = 8
    E, V = [], []
    with quiet():
        h.setDiagConfig(eigvalues=k, get_vectors=True)
        for c in grid[:, 0]:
            h.setParameterValue("C", float(c))
            e, v = h.diagonalize(h.getHamiltonian())
            E.append(np.asarray(e, dtype=np.float64))
            V.append(np.stack([x.data.to_array()[:, 0] for x in v], axis=1))
    out = {"E": np.array(E), "V": np.array(V)[[0, 4, 8]], "vec_points": np.array([0, 4, 8]), "params": grid,
           "swept_names": np.array(plan.swept_names), "k": np.int64(k), "n": np.int64(h.getHilbertSpaceSize()),
           "ref_seconds_per_point": np.float64(0.0)}
    np.savez_compressed(os.path.join(HERE, "c7_fluxonium_C.npz"), **out)
    plan.meta["source"] = "c7_fluxonium_C"
    plan.save(os.path.join(HERE, "c7_fluxonium_C.plan.json"))
    print("c7_fluxonium_C               n=%d pts=%d" % (int(out["n"]), len(grid)), flush=True)


ALL = {"big": make_big, "a4": make_a4, "single": make_single, "c6": make_c6, "c7": make_c7, "c2b": make_c2b, "c1": make_c1, "c2": make_c2, "c3": make_c3, "c4": make_c4, "c5": make_c5}

if __name__ == "__main__":
    which = sys.argv[1:] or list(ALL)
    for w in which:
        ALL[w]()
