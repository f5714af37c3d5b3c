"""Drop-in for the sweep / diagonalisation API of ``pyscqed.NumericalSystem``
(pyscqed/numerical_system.py:22-1287), backed by the B200 engine.

Same method names, argument meaning and error behaviour as the reference for the hot path:

    hamil = NumericalSystem(symbolic_system)            # :33
    hamil.configureOperator(node, trunc, basis)         # :133
    hamil.setParameterValues(...)                       # :831
    H = hamil.getHamiltonian(); H.eigenenergies()       # :509
    hamil.setDiagConfig(eigvalues=k, get_vectors=...)   # :790
    hamil.addSweep(name, start, end, N)                 # :864
    hamil.addEvaluation('Resonator', cpl_node=1)        # :868
    sweep = hamil.paramSweep()                          # :885
    x, Y, static = hamil.getSweep(sweep, 'phiZ', {})    # :879

The symbolic layers (``CircuitGraph``, ``SymbolicSystem``, ``ParamCollection``, ``Units``) are the
reference's own objects and are used as-is; everything numerical -- coefficient evaluation,
Hamiltonian assembly, eigensolves, resonator strips -- runs on the GPU through ``libscqed.so``.
There is no CPU fallback.

Differences, all deliberate:
  * ``paramSweep`` makes ONE engine call for the whole grid and returns an in-memory
    :class:`SweepData` instead of a list of per-point pickle paths (the reference writes one temp
    file per point, dataspec.py:83-104).  ``getSweep`` accepts it and returns arrays with exactly
    the reference's layout (parameters.py:1060-1068,1146-1157); ``SweepData.records()`` gives the
    reference's per-point records for code that indexed the old list.
  * eigenvectors are :class:`Ket` objects (``.full()``, ``.dims``, ``.norm()``, ``.dag_dot``);
    ``Ket.to_qobj()`` converts when QuTiP is installed.
  * oscillator nodes whose impedance depends on a swept parameter (per-point operator
    regeneration, :378-393) are refused with ``LoweringError`` for now.
"""
from __future__ import annotations

import time

import numpy as np
import sympy as sy

from . import engine as _engine
from .lowering import lower_system

_qobj_atol = 1e-12


class Ket:
    """Eigenvector with the tensor ``dims`` the reference attaches (util.py:167)."""
    __array_priority__ = 1000

    def __init__(self, data, dims):
        self._v = np.asarray(data, dtype=np.complex128).reshape(-1)
        self.dims = dims
        self.shape = (self._v.size, 1)

    def full(self):
        return self._v.reshape(-1, 1)

    def norm(self):
        return float(np.linalg.norm(self._v))

    def overlap(self, other):
        return complex(np.vdot(self._v, other._v))

    def to_qobj(self):
        import qutip as qt
        return qt.Qobj(self._v.reshape(-1, 1), dims=self.dims)

    def __truediv__(self, s):
        return Ket(self._v / s, self.dims)

    def __repr__(self):
        return "Ket(dims=%r, norm=%.6f)" % (self.dims, self.norm())


class _Data:
    def __init__(self, a):
        self._a = a

    def to_array(self):
        return self._a


class HamiltonianMatrix:
    """What ``getHamiltonian()`` returns: the dense H assembled on the GPU, with the handful of
    ``Qobj`` members the reference's users touch (examples/*.py: ``.eigenenergies()``,
    ``.data.to_array()``, ``.dims``, ``.shape``)."""

    def __init__(self, array, dims):
        self._a = np.asarray(array, dtype=np.complex128)
        self.dims = [list(dims), list(dims)]
        self.shape = self._a.shape
        self.data = _Data(self._a)

    def full(self):
        return self._a

    @property
    def isherm(self):
        return bool(np.max(np.abs(self._a - self._a.conj().T)) <= 1e-12)

    def eigenenergies(self):
        import torch
        H = torch.as_tensor(self._a[None].copy(), device="cuda")
        E, _, info = _engine.eigh_batched(H, self._a.shape[0], want_vectors=False)
        return E.cpu().numpy()[0]

    def to_qobj(self):
        import qutip as qt
        return qt.Qobj(self._a, dims=self.dims)


class SweepData:
    """In-memory result of ``paramSweep`` (all points, all evaluables)."""

    def __init__(self, specs, evaluations, E, V, scalars, Erwa, info, dims, timings, matel=None):
        self.specs = specs                 # [{'name','start','end','N'}]
        self.evaluations = evaluations     # [method names in addEvaluation order]
        self.E, self.V, self.scalars, self.Erwa, self.info = E, V, scalars, Erwa, info
        self.matel = matel or {}           # {'getCurrentMatrixElement': [n_pts, n_elements] float64, ...}
        self.dims = dims
        self.timings = timings

    def __len__(self):
        return len(self.E) if self.E is not None else len(next(iter(self.scalars.values())))

    def _kets(self, i):
        out = np.empty(self.V.shape[1], dtype=object)
        d = [list(self.dims), [1] * len(self.dims)]
        for j in range(self.V.shape[1]):
            out[j] = Ket(self.V[i, j], d)
        return out

    def record(self, i, key=None):
        """Per-point record in the reference's format (numerical_system.py:950-964,1021-1037)."""
        def one(name):
            if name == "getHamiltonian":
                if self.V is not None:
                    return (tuple(self.E[i].tolist()), self._kets(i))
                return self.E[i]
            if name == "getResonatorResponse":
                return self.Erwa[i]
            if name in self.matel:
                return self.matel[name][i]
            return self.scalars[name][i]
        if len(self.evaluations) == 1 and key is None:
            return one(self.evaluations[0])
        if key is not None:
            return one(key)
        return {name: one(name) for name in self.evaluations}

    def records(self):
        return [self.record(i) for i in range(len(self))]


_EVAL_SPEC = {   # numerical_system.py:499-507
    "Hamiltonian": {"eval": "getHamiltonian", "diag": True, "depends": None},
    "Resonator": {"eval": "getResonatorResponse", "diag": False, "depends": "getHamiltonian"},
    "Current": {"eval": "getCurrentMatrixElement", "diag": False, "depends": "getHamiltonian"},
    "Voltage": {"eval": "getVoltageMatrixElement", "diag": False, "depends": "getHamiltonian"},
    "ChargingEnergy": {"eval": "getChargingEnergies", "diag": False, "depends": None},
    "FluxEnergy": {"eval": "getFluxEnergies", "diag": False, "depends": None},
    "JosephsonEnergy": {"eval": "getJosephsonEnergies", "diag": False, "depends": None},
}


class NumericalSystem:
    def __init__(self, symbolic_system, unit=None, device=None):
        if unit is None:
            from pyscqed import units as _units      # the reference's unit presets, used as-is
            unit = _units.Units("CQED1")
        self.SS = symbolic_system
        self.units = unit
        self.device = device
        self.operator_data = {}
        #: swept oscillator impedance: 'consistent' = H(p) is what setParameterValues + getHamiltonian give at p;
        #: 'reference' = reproduce the reference's sweep loop, which keeps the charge / flux operators of the node
        #: at the pre-sweep impedance (numerical_system.py:259-271 are not re-run by _postsub)
        self.regen_mode = "consistent"
        self._circ_operators = None
        self._single = None            # (plan, engine) for the current parameter values
        self.setDiagConfig()
        self._init_sweep_data()
        self._set_parameter_units()
        self.getSymbolicExpressions()

    # ---- structure (numerical_system.py:64-100) -----------------------------------------------
    def getNodeList(self):
        return self.SS.nodes

    def getNodeIndex(self, node):
        return self.SS.nodes.index(node)

    def getEdgeList(self):
        return self.SS.edges

    def getEdgeIndex(self, edge):
        return self.SS.edges.index(edge)

    def getCircuitGraph(self):
        return self.SS.CG

    def getSymbolicSystem(self):
        return self.SS

    def getHilbertSpaceSize(self):
        ret = 1
        for k, v in self.operator_data.items():
            trunc, basis = v["truncation"], v["basis"]
            if basis == "charge":
                ret *= (2 * trunc + 1)
            elif basis == "oscillator":
                ret *= trunc
        return ret

    def _set_parameter_units(self):
        # numerical_system.py:1271-1287
        Uf = self.units.getUnitPrefactor('Hz')
        Uo = self.units.getUnitPrefactor('Ohm')
        Uc = self.units.getUnitPrefactor('F')
        Ul = self.units.getUnitPrefactor('H')
        if self.SS._has_resonators():
            for node, resonator in self.SS.CG.resonators_cap.items():
                if resonator is not None:
                    self.SS.addParameterisationPrefactor(resonator["Cr"], 1 / (Uf * Uo * Uc))
                    self.SS.addParameterisationPrefactor(resonator["Lr"], Uo / (Uf * Ul))
                    self.SS.addParameterisationPrefactor(resonator["gC"], self.units.getPrefactor('ChgOscCpl'))
                    self.SS.addParameterisationPrefactor(resonator["frl"], self.units.getPrefactor('Freq'))
                    self.SS.addParameterisationPrefactor(resonator["Zrl"], self.units.getPrefactor('Impe'))

    def getSymbolicExpressions(self):
        # numerical_system.py:280-297 (only what configureOperator needs is cached here)
        self.Cinv = self.SS.getInverseCapacitanceMatrix()
        self.Linv = self.SS.getInverseInductanceMatrix()

    # ---- operators (numerical_system.py:133-160) ------------------------------------------------
    def configureOperator(self, node, trunc, basis, fmax=4.0):
        if node not in self.getNodeList():
            raise Exception("Node '%i' is not a valid circuit node." % node)
        impedance = frequency = flux_max = None
        if basis == "oscillator":
            index = self.getNodeIndex(node)
            frequency = sy.sqrt(self.Linv[index, index] * self.Cinv[index, index])
            freq = "fosc%i" % node
            self.SS.addParameter(freq)
            self.SS.addParameterisation(freq, frequency)
            impedance = sy.sqrt(self.Cinv[index, index] / self.Linv[index, index])
            impe = "Zosc%i" % node
            self.SS.addParameter(impe)
            self.SS.addParameterisation(impe, impedance)
        elif basis == "discretized_flux":
            raise NotImplementedError("the reference's 'discretized_flux' basis is itself marked broken (numerical_system.py:179)")
        elif basis not in ("charge", "flux"):
            raise Exception("Unrecognized basis representation '%s'." % repr(basis))
        self.operator_data[node] = {"truncation": trunc, "basis": basis, "impedance": impedance,
                                    "frequency": frequency, "flux_max": flux_max}
        self._invalidate()

    @property
    def circ_operators(self):
        """``circ_operators[node][kind]``: full-space operators (numerical_system.py:185-253), built on
        demand as dense arrays -- an accessor for user code, not used by the engine."""
        if self._circ_operators is None:
            plan, _ = self._single_plan()
            dims = plan.dims
            out = {}
            for i, spec in enumerate(plan.nodes):
                ops = spec.operator_set()
                d = {}
                for kind in ("charge", "flux", "disp", "disp_adj", "pdisp", "pdisp_adj"):
                    m = np.eye(1, dtype=np.complex128)
                    for jn, dj in enumerate(dims):
                        m = np.kron(m, ops[kind] if jn == i else np.eye(dj))
                    d[kind] = HamiltonianMatrix(m, dims)
                out[spec.node] = d
            self._circ_operators = out
        return self._circ_operators

    # ---- parameters (numerical_system.py:817-854) -----------------------------------------------
    def _invalidate(self):
        if self._single is not None:
            self._single[1].close()
        self._single = None
        self._circ_operators = None

    def setParameterValue(self, name, value):
        self.SS.setParameterValue(name, value)
        params = self.getParameterValuesDict()
        assert all(v is not None for v in params.values()), \
            "Some parameters do not have valid values. All parameters should be set first with the " \
            "setParameterValues function in one go."
        self._invalidate()

    def getParameterValue(self, name):
        return self.SS.getParameterValue(name)

    def setParameterValues(self, *name_value_pairs):
        self.SS.setParameterValues(*name_value_pairs)
        params = self.getParameterValuesDict()
        assert all(v is not None for v in params.values()), "Not all parameters were set in this call. " \
            f"The missing parameters are {repr([n for n, v in params.items() if v is None])}"
        self._invalidate()

    def getParameterValues(self, *names):
        return self.SS.getParameterValues(*names)

    def getParameterValuesDict(self):
        return self.SS.getParameterValuesDict()

    def getParameterSweep(self, name):
        return self.SS.getParameterSweep(name)

    def getParameterNames(self):
        return self.SS.getParameterNamesList()

    def getPrefactor(self, name):
        return self.units.getPrefactor(name)

    # ---- single point ------------------------------------------------------------------------------
    def _scalar_exprs(self):
        out = {}
        Ec, El, Ej, Ep = (self.units.getPrefactor(x) for x in ("Ec", "El", "Ej", "Ep"))
        Cinv = self.SS.getInverseCapacitanceMatrix()
        Linv = self.SS.getInverseInductanceMatrix()
        J = self.SS.getJosephsonVector()
        V = self.SS.getPhaseSlipVector()
        for i, node in enumerate(self.SS.nodes):
            out["ChargingEnergy:%r" % (node,)] = sy.Rational(1, 2) * Cinv[i, i] * Ec     # :696-704
            out["FluxEnergy:%r" % (node,)] = sy.Rational(1, 2) * Linv[i, i] * El         # :706-714
        for i, edge in enumerate(self.SS.edges):
            out["JosephsonEnergy:%r" % (edge,)] = J[i] * Ej                              # :716-724
            out["PhaseSlipEnergy:%r" % (edge,)] = V[i] * Ep                              # :726-734
        return out

    def _single_plan(self):
        if self._single is None:
            plan = lower_system(self.SS, self.units, self.operator_data, scalars=self._scalar_exprs())
            self._single = (plan, _engine.Engine(plan, device=self.device))
        return self._single

    def getHamiltonian(self):
        """numerical_system.py:509-594, assembled on the device (K1 + K2)."""
        plan, eng = self._single_plan()
        H = eng.assemble_dense(np.zeros((1, 0)), tidyup=True).cpu().numpy()[0]
        return HamiltonianMatrix(H, plan.dims)

    def _single_scalars(self):
        plan, eng = self._single_plan()
        _, scal = eng.coefficients(np.zeros((1, 0)))
        return dict(zip(plan.program.scalar_names, scal.cpu().numpy()[0].tolist()))

    def _energy(self, prefix, keys, key):
        s = self._single_scalars()
        if key is None:
            return {k: s["%s:%r" % (prefix, k)] for k in keys}
        if key not in keys:
            raise ValueError("%r is not in list" % (key,))
        return s["%s:%r" % (prefix, key)]

    def getChargingEnergies(self, node=None):
        return self._energy("ChargingEnergy", self.getNodeList(), node)

    def getFluxEnergies(self, node=None):
        return self._energy("FluxEnergy", self.getNodeList(), node)

    def getJosephsonEnergies(self, edge=None):
        return self._energy("JosephsonEnergy", self.SS.edges, edge)

    def getPhaseSlipEnergies(self, edge=None):
        return self._energy("PhaseSlipEnergy", self.SS.edges, edge)

    # ---- Current / Voltage operators (numerical_system.py:596-694) --------------------------------
    def _aux_operator(self, spec):
        """Dense operator of one Current / Voltage spec at the current parameter values, assembled
        on the device by the same K1 + K2 kernels as H (its terms become the plan's term table)."""
        plan = lower_system(self.SS, self.units, self.operator_data, matrix_elements=[dict(spec, elements=[])])
        with _engine.Engine(plan.aux_as_terms(0), device=self.device) as eng:
            M = eng.assemble_dense(np.zeros((1, 0)), tidyup=False).cpu().numpy()[0]
        return HamiltonianMatrix(M, plan.dims)

    def getCurrentOperator(self, edge=None):
        return self._aux_operator({"kind": "Current", "edge": edge})

    def getVoltageOperator(self, node=None):
        return self._aux_operator({"kind": "Voltage", "node": node})

    def _aux_matrix_elements(self, spec, V):
        plan = lower_system(self.SS, self.units, self.operator_data, matrix_elements=[spec])
        vecs = np.stack([np.asarray(v.full() if hasattr(v, "full") else v, dtype=np.complex128).reshape(-1) for v in V])
        with _engine.Engine(plan, device=self.device) as eng:
            M = eng.matrix_elements(np.zeros((1, 0)), np.ascontiguousarray(vecs[None])).cpu().numpy()[0]
        return np.ascontiguousarray(M.real)

    def getCurrentMatrixElement(self, E, V, edge=None, elements=None):
        return self._aux_matrix_elements({"kind": "Current", "edge": edge, "elements": elements}, V)

    def getVoltageMatrixElement(self, E, V, node=None, elements=None):
        return self._aux_matrix_elements({"kind": "Voltage", "node": node, "elements": elements}, V)

    # ---- small conveniences of the reference API ---------------------------------------------------
    def prepareOperators(self):
        """numerical_system.py:302-305.  Operators are device tables built from the plan; nothing to expand."""
        self._invalidate()

    def sparsity(self, op):
        """numerical_system.py:99-100: 1 - nnz / n^2."""
        a = op.full() if hasattr(op, "full") else np.asarray(op.data.to_array())
        return 1 - np.count_nonzero(a) / self.getHilbertSpaceSize() ** 2

    def getOperatorList(self, node):
        """(Q, P, D, Ddag, S, Sdag) of one node as dense arrays (numerical_system.py:162-182)."""
        if node not in self.getNodeList():
            raise Exception("Node '%i' is not a valid circuit node." % node)
        plan, _ = self._single_plan()
        ops = plan.nodes[self.getNodeIndex(node)].operator_set()
        return tuple(ops[kind] for kind in ("charge", "flux", "disp", "disp_adj", "pdisp", "pdisp_adj"))

    def getResonatorResponse(self, E, V, nmax=100, cpl_node=None):
        """numerical_system.py:736-784 at the current parameter values.  The strips are computed on the device from
        the eigenpairs of H at these values (``E`` / ``V`` only fix how many levels enter, as in a sweep)."""
        if cpl_node is None:
            return None
        k = len(E)
        plan = lower_system(self.SS, self.units, self.operator_data, resonator={"cpl_node": cpl_node, "nmax": nmax})
        with _engine.Engine(plan, device=self.device) as eng:
            res = eng.sweep(np.zeros((1, 0)), k, want_vectors=True, resonator=True)
        return res.Erwa[0]

    # ---- diagonaliser (numerical_system.py:790-810; util.py:76-176) -----------------------------
    def setDiagConfig(self, eigvalues=5, get_vectors=False, sparse=False,
                      sparsesolveropts={"sigma": None, "mode": "normal", "maxiter": None, "tol": 1e-3, "which": "SA"}):
        # `sparse` / `sparsesolveropts` select ARPACK in the reference; the engine always returns the
        # exact lowest-k (dense-quality) eigenpairs, so they are recorded but do not change results.
        self.diagonalizer_config = {
            'kwargs': {'eigvalues': eigvalues, 'get_vectors': get_vectors, 'sparsesolveropts': sparsesolveropts},
            'sparse': sparse,
            'func': self._diag_device,
        }

    def getDiagConfig(self):
        return self.diagonalizer_config

    def _diag_device(self, M, eigvalues=5, get_vectors=False, sparsesolveropts=None):
        import torch
        if isinstance(M, HamiltonianMatrix):
            a, dims = M.full(), M.dims[0]
        elif hasattr(M, "data") and hasattr(M.data, "to_array"):       # a qutip.Qobj
            a, dims = np.asarray(M.data.to_array(), dtype=np.complex128), list(M.dims[0])
        else:
            raise Exception("not a qutip Qobj instance.")              # util.py:96-97,147-148
        H = torch.as_tensor(np.ascontiguousarray(a)[None].copy(), device="cuda")
        E, V, info = _engine.eigh_batched(H, int(eigvalues), want_vectors=bool(get_vectors))
        E = E.cpu().numpy()[0]
        if not get_vectors:
            return E
        V = V.cpu().numpy()[0]
        kets = np.empty(len(E), dtype=object)
        for j in range(len(E)):
            kets[j] = Ket(V[j], [list(dims), [1] * len(dims)])
        return tuple(E.tolist()), kets

    def diagonalize(self, M):
        return self.diagonalizer_config['func'](M, **self.diagonalizer_config['kwargs'])

    # ---- sweeps (numerical_system.py:860-928) -----------------------------------------------------
    def _init_sweep_data(self):
        self.sweep_specs = []
        self.evaluations = []

    def newSweep(self):
        self._init_sweep_data()

    def addSweep(self, *args, **kwargs):
        self.sweep_specs.append(self.SS.paramSweepSpec(*args, **kwargs))

    def addEvaluation(self, evaluable, **kwargs):
        if evaluable not in _EVAL_SPEC:
            raise Exception("Evaluable '%s' not valid." % evaluable)
        data = dict(_EVAL_SPEC[evaluable])
        data['kwargs'] = kwargs
        data['name'] = evaluable
        self.evaluations.append(data)

    def paramSweep(self, timesweep=True, keep_on_device=False):
        init_time = time.time()
        if len(self.evaluations) == 0:
            self.evaluations = [dict(_EVAL_SPEC["Hamiltonian"], kwargs={}, name="Hamiltonian")]
        self.SS.ndSweep(self.sweep_specs)                      # bounds checked here (parameters.py:182-189)
        specs = [dict(s) for s in self.sweep_specs]
        names = [s["name"] for s in specs]
        cfg = self.diagonalizer_config['kwargs']
        k, get_vectors = int(cfg['eigvalues']), bool(cfg['get_vectors'])

        want_h = False
        resonator = None
        scalars = {}
        matel_specs, matel_methods = [], []
        all_scalars = None
        methods = []
        for entry in self.evaluations:
            nm, kw = entry['name'], entry['kwargs']
            methods.append(entry['eval'])
            if entry['depends'] is not None:
                if 'getHamiltonian' not in methods[:-1]:
                    raise Exception("eval spec with 'depends':'%s' entry should be specified after the one it depends on ('%s'), or Possibly invalid 'depends' value." % (entry['depends'], entry['eval']))
                if not get_vectors:
                    raise Exception("need eigenvectors for 'depends'")          # :1009-1012
            if nm == "Hamiltonian":
                want_h = True
            elif nm == "Resonator":
                if kw.get("cpl_node") is None:
                    raise NotImplementedError("Resonator evaluation without cpl_node returns None in the reference")
                resonator = {"cpl_node": kw["cpl_node"], "nmax": kw.get("nmax", 100)}
            elif nm in ("ChargingEnergy", "FluxEnergy", "JosephsonEnergy"):
                if all_scalars is None:
                    all_scalars = self._scalar_exprs()
                key = kw.get("node") if nm != "JosephsonEnergy" else kw.get("edge")
                if key is None:
                    raise NotImplementedError("%s over all nodes/edges in a sweep: pass node=/edge=" % nm)
                scalars[entry['eval']] = all_scalars["%s:%r" % (nm, key)]
            elif nm in ("Current", "Voltage"):
                spec = dict(kw, kind=nm)
                spec.setdefault("elements", None)
                matel_specs.append(spec)
                matel_methods.append(entry['eval'])
            else:
                raise NotImplementedError("evaluable '%s' is not implemented on the GPU path yet" % nm)

        plan = lower_system(self.SS, self.units, self.operator_data, names, specs, scalars=scalars, resonator=resonator,
                            matrix_elements=matel_specs, regen_mode=self.regen_mode)
        grid = plan.sweep_grid()
        loop_time = time.time()
        matel = {}
        with _engine.Engine(plan, device=self.device) as eng:
            if want_h or resonator is not None:
                res = eng.sweep(grid, k, want_vectors=get_vectors, resonator=resonator is not None,
                                matrix_elements=bool(matel_specs), real_vectors="auto")
                E, V, scal, Erwa, info = res.E, res.V, res.scalars, res.Erwa, res.info
                col = 0
                for method, op in zip(matel_methods, plan.aux_ops):      # :656-667,683-694 keep the real part
                    ne = len(op["elements"])
                    matel[method] = np.ascontiguousarray(res.matel[:, col:col + ne].real)
                    col += ne
            else:
                _, sc = eng.coefficients(grid)
                sc = sc.cpu().numpy()
                scal = {nm: sc[:, i] for i, nm in enumerate(plan.program.scalar_names)}
                E = V = Erwa = None
                info = np.zeros(len(grid), dtype=np.int32)
        end_time = time.time()
        if info.any():
            raise RuntimeError("eigensolver did not converge at %d sweep point(s)" % int(np.count_nonzero(info)))
        data = SweepData(specs, methods, E, V, scal, Erwa, info, plan.dims,
                         {"init": loop_time - init_time, "loop": end_time - loop_time}, matel)
        self._init_sweep_data()                                   # :915
        if timesweep:
            print("Parameter Sweep Duration:")
            print("  Initialization:\t%.3f s" % (loop_time - init_time))
            print("  Loop duration:\t%.3f s" % (end_time - loop_time))
            print("  Avg iteration:\t%.3g s" % ((end_time - loop_time) / len(grid)))
        return data

    def getSweep(self, data, ind_var, static_vars, evaluable="Hamiltonian"):
        if evaluable not in _EVAL_SPEC:
            raise Exception("Evaluable '%s' not valid." % evaluable)
        key = _EVAL_SPEC[evaluable]['eval']
        if not isinstance(data, SweepData):
            return self.SS.getSweepResult(ind_var, static_vars, data=data, key=key)
        return _sweep_result(self.SS, data, ind_var, static_vars, key)


def _sweep_result(SS, data, ind_var, static_vars, key):
    """``ParamCollection.getSweepResult`` (parameters.py:1002-1166) on in-memory arrays: identical
    index selection and output layout (eigenvalue index first: ``np.array(records).T``)."""
    specs = data.specs
    if key not in data.evaluations:
        raise KeyError(key)
    names = [s["name"] for s in specs]
    Ns = {s["name"]: int(s["N"]) for s in specs}
    axes = {s["name"]: np.linspace(s["start"], s["end"], int(s["N"])) for s in specs}
    shape = [Ns[n] for n in names]

    def take(flat_idx):
        recs = [data.record(int(i), key) for i in flat_idx]
        if data.V is not None and key == "getHamiltonian":
            arr = np.empty((len(recs), 2, data.E.shape[1]), dtype=object)
            for a, (E, V) in enumerate(recs):
                arr[a, 0, :] = E
                arr[a, 1, :] = V
            return arr
        return np.array(recs)

    def nearest(static_vars):
        idx, vals = {}, {}
        for k_, v in static_vars.items():
            if k_ not in SS.getParameterNamesList():
                raise ValueError("Static variable '%s' does not exist." % k_)
            i = int(np.argmin(np.abs(axes[k_] - v)))
            idx[k_], vals[k_] = i, axes[k_][i]
        return idx, vals

    if type(ind_var) is str:
        if ind_var not in SS.getParameterNamesList():
            raise ValueError("Independent variable '%s' does not exist." % ind_var)
        if len(specs) == 1:
            return axes[ind_var], take(range(len(data))).T, {}
        sidx, svals = nearest(static_vars)
        multi = []
        for n in names:
            if n == ind_var:
                multi.append(np.arange(Ns[n]))
            else:
                multi.append(np.full(Ns[ind_var], sidx[n]))
        flat = np.ravel_multi_index(multi, shape)
        return axes[ind_var], take(flat).T, svals
    if type(ind_var) in (list, np.ndarray):
        for k_ in ind_var:
            if k_ not in SS.getParameterNamesList():
                raise ValueError("Independent variable '%s' does not exist." % k_)
        ordered = [n for n in names if n in ind_var]
        sidx, svals = nearest(static_vars)
        grids = np.meshgrid(*[np.arange(Ns[n]) if n in ordered else np.array([sidx[n]]) for n in names], indexing="ij")
        flat = np.ravel_multi_index([g.reshape(-1) for g in grids], shape)
        arr = take(flat)
        out = arr.reshape(*[Ns[n] for n in ordered], *arr.shape[1:]).T
        return [axes[n] for n in ordered], out, svals
    raise TypeError("Invalid independent variable specification. Found type '%s'" % repr(type(ind_var)))
