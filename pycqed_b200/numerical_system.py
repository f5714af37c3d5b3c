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
  * ``getSweep`` on a sweep with eigenvectors returns a lazy :class:`EVResult` that indexes like the
    reference's ``(k, 2, N)`` object array (``EV[:, 0]`` energies, ``EV[:, 1]`` kets); ``np.asarray(EV)``
    materialises the object array.  Eigenvectors stay on the GPU until something reads them.
  * oscillator nodes whose impedance depends on a swept parameter (per-point operator regeneration,
    :378-393) are lowered in the eigenbasis of a + a^dagger; ``regen_mode`` = 'reference' (default:
    what the reference's sweep loop computes, charge / flux operators left at the pre-sweep impedance)
    or 'consistent' (every operator follows the impedance).
  * lowered plans and uploaded engines are cached per circuit configuration;
    ``device=None`` uses every visible GPU of the box (one host thread per GPU, points sharded in
    contiguous blocks), ``device=i`` or a list restricts it.
"""
from __future__ import annotations

import time
from collections import OrderedDict
from dataclasses import replace as _dc_replace

import numpy as np
import sympy as sy

from . import engine as _engine
from .lowering import lower_system

_qobj_atol = 1e-12


class Ket:
    """Eigenvector with the tensor ``dims`` the reference attaches (util.py:167).  Holds a view of the sweep's
    eigenvector array; nothing is copied until ``full()`` is called."""
    __array_priority__ = 1000
    __slots__ = ("_v", "dims")

    def __init__(self, data, dims):
        self._v = data
        self.dims = dims

    @property
    def shape(self):
        return (int(np.size(self._v)), 1)

    def _vec(self):
        return np.asarray(self._v, dtype=np.complex128).reshape(-1)

    def full(self):
        return self._vec().reshape(-1, 1)

    def norm(self):
        return float(np.linalg.norm(self._vec()))

    def overlap(self, other):
        return complex(np.vdot(self._vec(), other._vec()))

    def to_qobj(self):
        import qutip as qt
        return qt.Qobj(self.full(), dims=self.dims)

    def __truediv__(self, s):
        return Ket(self._vec() / s, self.dims)

    def __repr__(self):
        return "Ket(dims=%r, norm=%.6f)" % (self.dims, self.norm())


class KetGrid:
    """Lazy array of kets: ``grid[idx]`` indexes like the numpy object array the reference returns, a scalar index
    gives a :class:`Ket`.  ``pt`` / ``lv`` hold, for every position, the sweep point and the level of the ket."""

    def __init__(self, source, pt, lv, dims):
        self._src = source            # callable -> eigenvector array [n_pts, k, n] (fetched from the GPU on first use)
        self._pt, self._lv, self._dims = pt, lv, dims
        self.shape = pt.shape
        self.ndim = pt.ndim
        self.dtype = np.dtype(object)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        pt, lv = self._pt[idx], self._lv[idx]
        if np.ndim(pt) == 0:
            return Ket(self._src()[int(pt), int(lv)], self._dims)
        return KetGrid(self._src, pt, lv, self._dims)

    def __iter__(self):
        return (self[i] for i in range(self.shape[0]))

    @property
    def T(self):
        return KetGrid(self._src, self._pt.T, self._lv.T, self._dims)

    def vectors(self):
        """The eigenvectors themselves, complex/float array of shape ``self.shape + (n,)`` (no per-ket objects)."""
        return self._src()[self._pt, self._lv]

    def __array__(self, dtype=None, copy=None):
        V = self._src()
        out = np.empty(self.shape, dtype=object)
        flat = out.reshape(-1)
        for a, (p, l) in enumerate(zip(self._pt.reshape(-1).tolist(), self._lv.reshape(-1).tolist())):
            flat[a] = Ket(V[p, l], self._dims)
        return out


class EVResult:
    """What ``getSweep`` returns for a sweep with eigenvectors: indexes like the reference's object array of shape
    ``(k, 2, ...)`` (parameters.py:1063-1068): ``EV[:, 0]`` -> energies (float array), ``EV[:, 1]`` -> :class:`KetGrid`.
    Any other indexing pattern goes through the materialised object array (``np.asarray(EV)``)."""

    def __init__(self, E, kets):
        self._E, self._kets = E, kets              # both of shape (k, ...)
        self.shape = (E.shape[0], 2) + tuple(E.shape[1:])
        self.ndim = len(self.shape)
        self.dtype = np.dtype(object)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, tuple) and len(idx) >= 2 and isinstance(idx[1], (int, np.integer)) and not any(i is Ellipsis for i in idx[:2]):
            which = int(idx[1])
            if which not in (0, 1, -1, -2):
                raise IndexError("index %d is out of bounds for axis 1 with size 2" % which)
            rest = (idx[0],) + tuple(idx[2:])
            return (self._E if which in (0, -2) else self._kets)[rest]
        return np.asarray(self)[idx]

    def __array__(self, dtype=None, copy=None):
        out = np.empty(self.shape, dtype=object)
        out[:, 0] = self._E
        out[:, 1] = np.asarray(self._kets)
        return out


class _Data:
    def __init__(self, a):
        self._a = a

    def to_array(self):
        return self._a


class HamiltonianMatrix:
    """What ``getHamiltonian()`` returns: the dense H assembled on the GPU, with the handful of
    ``Qobj`` members the reference's users touch (examples/*.py: ``.eigenenergies()``,
    ``.data.to_array()``, ``.dims``, ``.shape``)."""

    def __init__(self, array, dims, device=None):
        self._a = np.asarray(array, dtype=np.complex128)
        self._device = device
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
        dev = torch.device("cuda", self._device if self._device is not None else torch.cuda.current_device())
        H = torch.as_tensor(self._a[None].copy(), device=dev)
        E, _, info = _engine.eigh_batched(H, self._a.shape[0], want_vectors=False)
        if info.any():
            raise RuntimeError("eigensolver did not converge")
        return E.cpu().numpy()[0]

    def to_qobj(self):
        import qutip as qt
        return qt.Qobj(self._a, dims=self.dims)


class SweepData:
    """In-memory result of ``paramSweep`` (all points, all evaluables).  Arrays: ``E`` [n_pts, k], ``Erwa``
    [n_pts, nmax+1+k, k], ``scalars`` {name: [n_pts]}, ``matel`` {method: [n_pts, n_elements]}, ``info`` [n_pts];
    ``V`` [n_pts, k, n] is fetched from the GPU(s) the first time it is read."""

    def __init__(self, specs, evaluations, res, scalars, dims, timings, matel=None, scalar_map=None, none_methods=()):
        self.specs = specs                 # [{'name','start','end','N'}]
        self.evaluations = evaluations     # [method names in addEvaluation order]
        self._res = res                    # engine.SweepResult or None (scalar-only sweeps)
        self.E = res.E if res is not None else None
        self.Erwa = res.Erwa if res is not None else None
        self.info = res.info if res is not None else None
        self.scalars = scalars
        self.matel = matel or {}           # {'getCurrentMatrixElement': [n_pts, n_elements] float64, ...}
        self.scalar_map = scalar_map or {} # method -> scalar name, or {node / edge: scalar name} (all-nodes form)
        self.none_methods = set(none_methods)
        self.dims = dims
        self.timings = timings

    @property
    def has_vectors(self):
        return self._res is not None and (self._res.V is not None or bool(self._res.V_dev))

    @property
    def V(self):
        return self._res.vectors() if self._res is not None else None

    def __len__(self):
        return len(self.E) if self.E is not None else len(next(iter(self.scalars.values())))

    def ket_dims(self):
        return [list(self.dims), [1] * len(self.dims)]

    def _kets(self, i):
        V, d = self.V, self.ket_dims()
        out = np.empty(V.shape[1], dtype=object)
        for j in range(V.shape[1]):
            out[j] = Ket(V[i, j], d)
        return out

    def record(self, i, key=None):
        """Per-point record in the reference's format (numerical_system.py:950-964,1021-1037)."""
        def one(name):
            if name == "getHamiltonian":
                if self.has_vectors:
                    return (tuple(self.E[i].tolist()), self._kets(i))
                return self.E[i]
            if name in self.none_methods:
                return None
            if name == "getResonatorResponse":
                return self.Erwa[i]
            if name in self.matel:
                return self.matel[name][i]
            m = self.scalar_map.get(name, name)
            if isinstance(m, dict):                      # all nodes / edges: {node: value} (:696-724)
                return {k_: self.scalars[nm][i] for k_, nm in m.items()}
            return self.scalars[m][i]
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
        #: GPUs used: None = every visible GPU of the box (sweep points sharded in contiguous blocks, one host thread
        #: per GPU), an int = that device, a list = those devices
        self.device = device
        self.operator_data = {}
        #: swept oscillator impedance: 'reference' (default, drop-in parity) = reproduce the reference's sweep loop, which
        #: keeps the charge / flux operators of the node at the pre-sweep impedance (numerical_system.py:259-271 are not
        #: re-run by _postsub); 'consistent' = H(p) is what setParameterValues + getHamiltonian give at every point p
        self.regen_mode = "reference"
        #: paramSweep leaves eigenvectors on the GPU and fetches them when first read ("auto": when they fit a quarter
        #: of the free device memory; True / False force it)
        self.lazy_vectors = "auto"
        self._circ_operators = None
        self._cache = OrderedDict()    # plan / engine cache, see _lowered()
        self._cache_max = 8
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

    # ---- plan / engine cache -----------------------------------------
    def _devices(self):
        if self.device is None:
            devs = _engine.visible_devices()
            if not devs:
                raise _engine.EngineError("no CUDA device visible; the sweep engine has no CPU fallback")
            return devs
        if isinstance(self.device, (list, tuple)):
            return [int(d) for d in self.device]
        return [int(self.device)]

    def _config_key(self, swept):
        """Everything a lowered plan depends on apart from the request itself: operator configuration, the
        parameterisations and the values of every parameter that is not swept."""
        SS = self.SS
        ops = tuple(sorted((repr(n), int(d["truncation"]), d["basis"]) for n, d in self.operator_data.items()))
        vals = tuple(sorted((nm, v) for nm, v in SS.getParameterValuesDict().items() if nm not in swept))
        par = tuple((nm, SS.getParametricExpression(nm)) for nm in sorted(SS.getParametricParametersList()))  # sympy expressions hash structurally (cached): no str()
        return ops, vals, par

    def _lowered(self, swept_names=(), sweep_specs=(), scalars=None, resonator=None, matrix_elements=(), regen_mode=None,
                 aux_as_terms=None):
        """(plan, EnginePool) for this request at the current circuit configuration.  ``lower_system`` (10-110 ms of
        sympy) and the upload of the tables happen once per distinct request; a later call -- another ``paramSweep``
        over a different grid of the same parameters, ``getHamiltonian`` again, a matrix-element evaluation -- reuses
        both."""
        regen_mode = regen_mode or self.regen_mode
        swept = tuple(swept_names)
        req = (swept, tuple(sorted(scalars)) if scalars else (), repr(sorted(resonator.items())) if resonator else None,
               repr([sorted((k_, repr(v)) for k_, v in m.items()) for m in matrix_elements]), regen_mode, aux_as_terms)
        # in 'reference' regen mode the pre-sweep value of a swept parameter enters the plan (stale impedance)
        key = (req, self._config_key(() if regen_mode == "reference" else swept), tuple(self._devices()))
        hit = self._cache.get(key)
        if hit is None:
            plan = lower_system(self.SS, self.units, self.operator_data, list(swept), list(sweep_specs), scalars=scalars,
                                resonator=resonator, matrix_elements=list(matrix_elements), regen_mode=regen_mode,
                                matrices={"Cinv": self.Cinv, "Linv": self.Linv})
            if aux_as_terms is not None:
                plan = plan.aux_as_terms(aux_as_terms)
            hit = [plan, None]
            self._cache[key] = hit
            while len(self._cache) > self._cache_max:
                _, (_, old) = self._cache.popitem(last=False)
                if old is not None:
                    old.close()
        else:
            self._cache.move_to_end(key)
        plan = hit[0]
        if sweep_specs:
            plan = _dc_replace(plan, sweep_specs=[dict(name=s_["name"], start=float(s_["start"]), end=float(s_["end"]),
                                                       N=int(s_["N"])) for s_ in sweep_specs])
        if hit[1] is None:
            hit[1] = _engine.EnginePool(hit[0], self._devices())
        return plan, hit[1]

    def close(self):
        """Free the device tables of every cached engine."""
        for _, pool in self._cache.values():
            if pool is not None:
                pool.close()
        self._cache.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters (numerical_system.py:817-854) -----------------------------------------------
    def _invalidate(self):
        # plans are keyed on the configuration they were lowered for: nothing to flush, only the lazily built
        # user-facing operator dictionary
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
        Cinv, Linv = self.Cinv, self.Linv          # cached at construction (each call re-derives a symbolic inverse)
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
        plan, pool = self._lowered(scalars=self._scalar_exprs())
        return plan, pool.engines[0]

    def getHamiltonian(self):
        """numerical_system.py:509-594, assembled on the device (K1 + K2)."""
        plan, eng = self._single_plan()
        H = eng.assemble_dense(np.zeros((1, 0)), tidyup=True).cpu().numpy()[0]
        return HamiltonianMatrix(H, plan.dims, device=getattr(getattr(eng, "device", None), "index", None))

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
        plan, pool = self._lowered(matrix_elements=[dict(spec, elements=[])], aux_as_terms=0)
        M = pool.engines[0].assemble_dense(np.zeros((1, 0)), tidyup=False).cpu().numpy()[0]
        return HamiltonianMatrix(M, plan.dims)

    def getCurrentOperator(self, edge=None):
        return self._aux_operator({"kind": "Current", "edge": edge})

    def getVoltageOperator(self, node=None):
        return self._aux_operator({"kind": "Voltage", "node": node})

    def _aux_matrix_elements(self, spec, V):
        spec = dict(spec)
        if spec.get("elements") is not None:
            spec["elements"] = [tuple(int(x) for x in e) for e in spec["elements"]]
        if isinstance(spec.get("edge"), list):
            spec["edge"] = tuple(spec["edge"])
        plan, pool = self._lowered(matrix_elements=[spec])
        if isinstance(V, KetGrid):
            vecs = np.asarray(V.vectors(), dtype=np.complex128)
        else:
            vecs = np.stack([np.asarray(v.full() if hasattr(v, "full") else v, dtype=np.complex128).reshape(-1) for v in V])
        M = pool.engines[0].matrix_elements(np.zeros((1, 0)), np.ascontiguousarray(vecs[None])).cpu().numpy()[0]
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
        plan, pool = self._lowered(resonator={"cpl_node": cpl_node, "nmax": int(nmax)})
        res = pool.engines[0].sweep(np.zeros((1, 0)), k, want_vectors=True, resonator=True)
        return np.array(res.Erwa[0])

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
        dev = torch.device("cuda", self._devices()[0])
        H = torch.as_tensor(np.ascontiguousarray(a)[None].copy(), device=dev)
        E, V, info = _engine.eigh_batched(H, int(eigvalues), want_vectors=bool(get_vectors))
        if info.any():
            raise RuntimeError("eigensolver did not converge")
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

    def paramSweep(self, timesweep=True):
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
        scalars, scalar_map, none_methods = {}, {}, []
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
                    none_methods.append(entry['eval'])                           # the reference returns None (:737-741)
                else:
                    resonator = {"cpl_node": kw["cpl_node"], "nmax": int(kw.get("nmax", 100))}
            elif nm in ("ChargingEnergy", "FluxEnergy", "JosephsonEnergy"):
                if all_scalars is None:
                    all_scalars = self._scalar_exprs()
                key = kw.get("node") if nm != "JosephsonEnergy" else kw.get("edge")
                keys = self.getNodeList() if nm != "JosephsonEnergy" else self.SS.edges
                if key is None:                                                  # all nodes / edges: a dict per point (:696-724)
                    scalar_map[entry['eval']] = {}
                    for kk in keys:
                        sname = "%s:%r" % (nm, kk)
                        scalars[sname] = all_scalars[sname]
                        scalar_map[entry['eval']][kk] = sname
                else:
                    if key not in keys:
                        raise ValueError("%r is not in list" % (key,))
                    sname = "%s:%r" % (nm, key)
                    scalars[sname] = all_scalars[sname]
                    scalar_map[entry['eval']] = sname
            elif nm in ("Current", "Voltage"):
                spec = dict(kw, kind=nm)
                spec.setdefault("elements", None)
                if spec["elements"] is not None:
                    spec["elements"] = [tuple(int(x) for x in e) for e in spec["elements"]]
                matel_specs.append(spec)
                matel_methods.append(entry['eval'])
            else:
                raise NotImplementedError("evaluable '%s' is not implemented on the GPU path yet" % nm)

        plan, pool = self._lowered(names, specs, scalars=scalars, resonator=resonator, matrix_elements=matel_specs)
        grid = plan.sweep_grid()
        loop_time = time.time()
        matel = {}
        res = None
        if want_h or resonator is not None:
            lazy = self.lazy_vectors
            if lazy == "auto":
                lazy = get_vectors and self._vectors_fit(len(grid), k, plan.hilbert_dim, len(pool.engines))
            res = pool.sweep(grid, k, want_vectors=get_vectors, resonator=resonator is not None,
                             matrix_elements=bool(matel_specs), real_vectors=False if lazy else "auto",
                             vectors_on_device=bool(lazy and get_vectors))
            scal, info = res.scalars, res.info
            col = 0
            for method, op in zip(matel_methods, plan.aux_ops):      # :656-667,683-694 keep the real part
                ne = len(op["elements"])
                matel[method] = np.ascontiguousarray(res.matel[:, col:col + ne].real)
                col += ne
        else:
            _, sc = pool.engines[0].coefficients(grid)
            sc = sc.cpu().numpy()
            scal = {nm: sc[:, i] for i, nm in enumerate(plan.program.scalar_names)}
            info = np.zeros(len(grid), dtype=np.int32)
        end_time = time.time()
        if info.any():
            raise RuntimeError("eigensolver did not converge at %d sweep point(s)" % int(np.count_nonzero(info)))
        data = SweepData(specs, methods, res, scal, plan.dims,
                         {"init": loop_time - init_time, "loop": end_time - loop_time}, matel, scalar_map, none_methods)
        self._init_sweep_data()                                   # :915
        if timesweep:
            print("Parameter Sweep Duration:")
            print("  Initialization:\t%.3f s" % (loop_time - init_time))
            print("  Loop duration:\t%.3f s" % (end_time - loop_time))
            print("  Avg iteration:\t%.3g s" % ((end_time - loop_time) / len(grid)))
        return data

    @staticmethod
    def _vectors_fit(n_pts, k, n, n_gpus):
        """Eigenvectors of one GPU's shard within a quarter of its free memory?"""
        try:
            import torch
            free, _ = torch.cuda.mem_get_info()
        except Exception:
            return False
        return (n_pts / max(1, n_gpus)) * k * n * 16 <= 0.25 * free

    def getSweep(self, data, ind_var, static_vars, evaluable="Hamiltonian"):
        if evaluable not in _EVAL_SPEC:
            raise Exception("Evaluable '%s' not valid." % evaluable)
        key = _EVAL_SPEC[evaluable]['eval']
        if not isinstance(data, SweepData):
            return self.SS.getSweepResult(ind_var, static_vars, data=data, key=key)
        return _sweep_result(self.SS, data, ind_var, static_vars, key)


def _sweep_result(SS, data, ind_var, static_vars, key):
    """``ParamCollection.getSweepResult`` (parameters.py:1002-1166) on in-memory arrays: identical
    index selection and output layout (eigenvalue index first: ``np.array(records).T``), built with array
    indexing -- no per-point Python objects (eigenvectors come back as a lazy :class:`EVResult`)."""
    specs = data.specs
    if key not in data.evaluations:
        raise KeyError(key)
    names = [s["name"] for s in specs]
    Ns = {s["name"]: int(s["N"]) for s in specs}
    axes = {s["name"]: np.linspace(s["start"], s["end"], int(s["N"])) for s in specs}
    shape = [Ns[n] for n in names]

    def take(flat_idx, lead=None):
        """records[flat_idx] stacked like np.array(records), optionally reshaped to ``lead`` + record shape,
        then transposed (.T) as the reference does."""
        flat = np.asarray(list(flat_idx) if isinstance(flat_idx, range) else flat_idx, dtype=np.int64)

        def fin(a):
            if lead is not None:
                a = a.reshape(*lead, *a.shape[1:])
            return a.T
        if key == "getHamiltonian":
            E = data.E[flat]
            if not data.has_vectors:
                return fin(E)
            k = E.shape[1]
            pt = np.broadcast_to(flat[:, None], (len(flat), k))
            lv = np.broadcast_to(np.arange(k)[None, :], (len(flat), k))
            # record = (E[k], V[k]) -> np.array(records) has shape (len, 2, k); .T puts the level first and the
            # (E, V) selector second: the two halves are transposed separately with the selector axis left out
            return EVResult(fin(E), KetGrid(lambda: data.V, fin(pt), fin(lv), data.ket_dims()))
        if key in data.none_methods:
            return fin(np.full(len(flat), None, dtype=object))
        if key == "getResonatorResponse":
            return fin(data.Erwa[flat])
        if key in data.matel:
            return fin(data.matel[key][flat])
        m = data.scalar_map.get(key, key)
        if isinstance(m, dict):
            arr = np.empty(len(flat), dtype=object)
            for a, i in enumerate(flat.tolist()):
                arr[a] = {k_: data.scalars[nm][i] for k_, nm in m.items()}
            return fin(arr)
        return fin(data.scalars[m][flat])

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
            return axes[ind_var], take(range(len(data))), {}
        sidx, svals = nearest(static_vars)
        multi = []
        for n in names:
            if n == ind_var:
                multi.append(np.arange(Ns[n]))
            else:
                multi.append(np.full(Ns[ind_var], sidx[n]))
        flat = np.ravel_multi_index(multi, shape)
        return axes[ind_var], take(flat), svals
    if type(ind_var) in (list, np.ndarray):
        for k_ in ind_var:
            if k_ not in SS.getParameterNamesList():
                raise ValueError("Independent variable '%s' does not exist." % k_)
        ordered = [n for n in names if n in ind_var]
        sidx, svals = nearest(static_vars)
        grids = np.meshgrid(*[np.arange(Ns[n]) if n in ordered else np.array([sidx[n]]) for n in names], indexing="ij")
        flat = np.ravel_multi_index([g.reshape(-1) for g in grids], shape)
        return [axes[n] for n in ordered], take(flat, lead=[Ns[n] for n in ordered]), svals
    raise TypeError("Invalid independent variable specification. Found type '%s'" % repr(type(ind_var)))
