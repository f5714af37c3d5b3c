"""Lowering: symbolic circuit + operator configuration + sweep spec -> :class:`SweepPlan`.

This is the once-per-sweep replacement of what the reference does at EVERY sweep point:
``_postsub`` (pyscqed/numerical_system.py:395-437) followed by ``getHamiltonian`` (:509-594).
The algebra of ``getHamiltonian`` is expanded symbolically here, once:

  Hq = 1/2 Ec (Q+q_b)^T C^-1 (Q+q_b)     -> terms  Q_i Q_j,  Q_i,  identity
  Hf = 1/2 El (P+phi_b)^T L^-1 (P+phi_b) -> terms  P_i P_j,  P_i,  identity
  Hj = -1/2 Ej sum_e J_e (e^{+2 pi i Phi_e} Pi_e^+ + e^{-2 pi i Phi_e} Pi_e^-)
  Hp = -1/2 Ep sum_e V_e (e^{+2 pi i q_e}  Sigma_e^+ + e^{-2 pi i q_e}  Sigma_e^-)

with Pi_e^+- the per-node products of D / D^dagger selected by the SIGN of the branch
coordinate coefficients exactly as numerical_system.py:529-552 does.

Oscillator nodes whose impedance depends on a swept parameter ("regen" nodes, which the reference
rebuilds at every point, numerical_system.py:378-393,435-437) are lowered in the eigenbasis of
a + a^dagger: their operators become constant unit matrices, the impedance dependence goes into the
coefficients, and exp(+-i beta(Z) X) becomes one projector term per level
(:func:`pycqed_b200.operators.oscillator_regen_set`).

Only public attributes/methods of the reference's ``SymbolicSystem`` are used (the symbolic
layer is reused, not rebuilt): ``nodes, edges, Rnb, Rinv, node_vector, node_map_rev,
cooper_disp, fluxon_disp, getInverse*Matrix, get*Vector, getFluxBiasMatrix,
getParameterFromSymbol, getParametricParametersList, getParametricExpression,
getParameterValue, getNonSweepParametersDict, getSymbolValuesDict``.
"""
from __future__ import annotations

import numpy as np
import sympy as sy

from .coefprog import ProgramBuilder
from .plan import NodeSpec, SweepPlan


class LoweringError(Exception):
    pass


def _resolve(expr, SS, swept_names, presub):
    """Rewrite ``expr`` in terms of the swept INDEPENDENT symbols only.

    After ``_presub`` (numerical_system.py:357-373) the reference's matrices still contain the
    swept symbols plus every parametric dependent of them (``getSweepParametersDict``,
    parameters.py:888-906); it fills those in numerically at each point through
    ``_update_parameterisations`` (parameters.py:1207-1240).  Here dependents are replaced by
    their defining expressions once, and everything not swept is a number.
    """
    expr = sy.sympify(expr)
    parametric = set(SS.getParametricParametersList())
    for _ in range(64):
        expr = expr.subs(presub)
        todo = {}
        for sym in expr.free_symbols:
            name = SS.getParameterFromSymbol(sym)
            if name in swept_names:
                continue
            if name in parametric:
                todo[sym] = SS.getParametricExpression(name)
            else:
                val = SS.getParameterValue(name)
                if val is None:
                    raise LoweringError("parameter '%s' has no value" % name)
                todo[sym] = val
        if not todo:
            return expr
        expr = expr.subs(todo)
    raise LoweringError("parameterisation expansion did not terminate for %r" % (expr,))


def _branch_composition(SS, i):
    """[(node, sign)] of branch i in node coordinates, parsed like numerical_system.py:519-552."""
    Pp = SS.Rnb * SS.Rinv * SS.node_vector
    e = Pp[i]
    out = []
    if len(e.atoms()) > 2:            # a sum of coef*symbol terms
        for arg in e.args:
            out.append((SS.node_map_rev[arg.args[1]], 1 if arg.args[0] > 0 else -1))
    else:
        out.append((SS.node_map_rev[e.args[1]], 1 if e.args[0] > 0 else -1))
    return out


def _expand_regen(regen, node_factor_names, re_expr, im_expr, fresh=False):
    """Rewrite a term that touches regenerated oscillator nodes: scale by q(Z) / f(Z) powers and expand
    a displacement operator into projector terms.  Yields (names, re, im).

    ``rg['stale']`` (regen_mode='reference'): the reference's sweep loop regenerates ``circ_operators`` but NOT
    the ``Qnp`` / ``Pnp`` vectors that ``getHamiltonian`` / ``getVoltageOperator`` / ``getResonatorResponse`` read
    (numerical_system.py:259-271 are only called from ``prepareOperators``, :302-305, not from ``_postsub``,
    :435-437), so Q and P stay at the impedance of the last ``setParameterValues`` while D follows the sweep.
    ``fresh`` marks the one consumer that does read the regenerated operators (``getCurrentOperator``, :608-620)."""
    names = dict(node_factor_names)
    scale = sy.Integer(1)
    disp = None
    for ni, name in node_factor_names.items():
        if ni not in regen:
            continue
        rg = regen[ni]
        q, f = (rg["q0"], rg["f0"]) if (rg["stale"] and not fresh) else (rg["q"], rg["f"])
        if name == "charge":
            scale *= q
        elif name == "charge2":
            scale *= q ** 2
        elif name == "flux":
            scale *= f
        elif name == "flux2":
            scale *= f ** 2
        elif name in ("disp", "disp_adj"):
            if disp is not None:
                raise LoweringError("a junction between two oscillator nodes whose impedances are both swept is not supported")
            disp = (ni, +1 if name == "disp" else -1)
        else:
            raise LoweringError("phase-slip displacement on an oscillator node with swept impedance is not supported")
    re_expr, im_expr = scale * re_expr, scale * im_expr
    if disp is None:
        yield names, re_expr, im_expr
        return
    ni, sgn = disp
    rg = regen[ni]
    for m, lam in enumerate(rg["lam"]):
        ph = rg["beta"] * float(lam)
        c, s_ = sy.cos(ph), sgn * sy.sin(ph)
        nm = dict(names)
        nm[ni] = "proj:%d" % m
        yield nm, re_expr * c - im_expr * s_, re_expr * s_ + im_expr * c


class _TermTable:
    def __init__(self, n_nodes, regen=None):
        self.n_nodes = n_nodes
        self.regen = regen or {}
        self.factors = []          # (node_index, name)
        self._fid = {}
        self.terms = {}            # tuple(factor ids) -> [re_expr, im_expr]
        self.order = []

    def fid(self, node_index, name):
        key = (node_index, name)
        if key not in self._fid:
            self._fid[key] = len(self.factors)
            self.factors.append(key)
        return self._fid[key]

    def add(self, node_factor_names: dict, re_expr, im_expr=0):
        """node_factor_names: {node_index: factor name}; coefficient = re + i im."""
        if re_expr == 0 and im_expr == 0:
            return
        if self.regen and any(ni in self.regen for ni in node_factor_names):
            expanded = list(_expand_regen(self.regen, node_factor_names, sy.sympify(re_expr), sy.sympify(im_expr)))
        else:
            expanded = [(node_factor_names, re_expr, im_expr)]
        for names, re_e, im_e in expanded:
            key = [-1] * self.n_nodes
            for ni, name in names.items():
                key[ni] = self.fid(ni, name)
            key = tuple(key)
            if key not in self.terms:
                self.terms[key] = [sy.Integer(0), sy.Integer(0)]
                self.order.append(key)
            self.terms[key][0] += re_e
            self.terms[key][1] += im_e


def lower_system(SS, units, operator_data, swept_names=(), sweep_specs=(), scalars=None,
                 resonator=None, matrix_elements=(), regen_mode="consistent", matrices=None) -> SweepPlan:
    """Build the plan.

    :param SS: the reference ``SymbolicSystem`` (after ``ndSweep`` if ``swept_names``)
    :param units: the reference ``Units`` object (``getPrefactor``)
    :param operator_data: ``{node: {'basis','truncation',...}}`` as kept by ``configureOperator``
    :param swept_names: independent parameters being swept, in ``addSweep`` order
    :param scalars: {name: sympy expr or parameter name} extra real per-point outputs
    :param resonator: {'cpl_node': node, 'nmax': int} to add the Resonator post-op inputs
    :param regen_mode: oscillator nodes with swept impedance: 'consistent' (every operator of the node follows
                       the impedance, i.e. H equals what ``setParameterValues`` + ``getHamiltonian`` give at each
                       point) or 'reference' (reproduce the reference's sweep loop, which leaves the charge /
                       flux operators at the pre-sweep impedance; see _expand_regen)
    :param matrices: optional {'Cinv': ..., 'Linv': ...}: the symbolic inverse capacitance / inductance matrices when
                     the caller already holds them (``SymbolicSystem.getInverse*Matrix`` re-derives a symbolic matrix
                     inverse and re-registers the parameterisations on every call, symbolic_system.py:194-224,363-391)
    :param matrix_elements: [{'kind': 'Voltage', 'node': n, 'elements': [(i, j), ...]} |
                             {'kind': 'Current', 'edge': e, 'elements': [...]}] (numerical_system.py:596-694)
    """
    swept_names = list(swept_names)
    nodes = list(SS.nodes)
    N = len(nodes)
    for node in nodes:
        if node not in operator_data:
            raise LoweringError("node %r has no operator configuration (configureOperator)" % (node,))

    if swept_names:
        presub = SS.getNonSweepParametersDict()
    else:
        presub = SS.getSymbolValuesDict()
    missing = [str(k) for k, v in presub.items() if v is None]
    if missing:
        raise LoweringError("parameters without a value: %s" % missing)
    swept_syms = [SS.getSymbol(n) for n in swept_names]

    def R(expr):
        return _resolve(expr, SS, swept_names, presub)

    # An oscillator node whose impedance depends on a swept symbol: the reference regenerates its
    # operators at every point (numerical_system.py:378-393,435-437); here the dependence is moved
    # into the coefficients (see _expand_regen).
    from . import operators as _ops
    regen = {}
    for node, data in operator_data.items():
        if data["basis"] == "oscillator":
            z = R(data["impedance"])
            if z.free_symbols:
                zi = z * units.getPrefactor("Impe")
                z0 = float(sy.sympify(data["impedance"]).subs(SS.getSymbolValuesDict())) * units.getPrefactor("Impe")
                if regen_mode not in ("consistent", "reference"):
                    raise LoweringError("regen_mode must be 'consistent' or 'reference'")
                regen[nodes.index(node)] = {
                    "stale": regen_mode == "reference",
                    "q0": float(np.sqrt(1.0 / (2.0 * z0))) * units.getPrefactor("ChgOsc"),
                    "f0": float(np.sqrt(z0 / 2.0)) * units.getPrefactor("FlxOsc"),
                    "q": sy.sqrt(1 / (2 * zi)) * units.getPrefactor("ChgOsc"),           # :1076
                    "f": sy.sqrt(zi / 2) * units.getPrefactor("FlxOsc"),                 # :1077
                    "beta": float(SS.cooper_disp[node]) * 2.0 * sy.pi / _ops.PHI0 * float(_ops.HBAR) ** 0.5 * sy.sqrt(zi / 2),   # :1078
                    "lam": _ops.oscillator_regen_set(int(data["truncation"]))["lam"],
                }

    Ec = units.getPrefactor("Ec")
    El = units.getPrefactor("El")
    Ej = units.getPrefactor("Ej")
    Ep = units.getPrefactor("Ep")

    matrices = matrices or {}
    Cinv = (matrices["Cinv"] if matrices.get("Cinv") is not None else SS.getInverseCapacitanceMatrix()).applyfunc(R)
    Linv = (matrices["Linv"] if matrices.get("Linv") is not None else SS.getInverseInductanceMatrix()).applyfunc(R)
    Jvec = SS.getJosephsonVector().applyfunc(R)
    Pvec = SS.getPhaseSlipVector().applyfunc(R)
    Qb = SS.getChargeBiasVector().applyfunc(R)
    Qbt = (SS.Rnb * SS.getChargeBiasVector()).applyfunc(R)
    Pbm = SS.getFluxBiasMatrix(mode="branch").applyfunc(R)
    Pbi = SS.getFluxBiasVectorInd().applyfunc(R)

    tt = _TermTable(N, regen)
    half = sy.Rational(1, 2)

    def quadratic(M, bias, pref, lin_name, sq_name):
        # pref/2 * sum_ij (O_i + b_i) M_ij (O_j + b_j)
        for i in range(N):
            for j in range(N):
                m = M[i, j]
                if m == 0:
                    continue
                c = half * pref * m
                if i == j:
                    tt.add({i: sq_name}, c)
                else:
                    tt.add({i: lin_name, j: lin_name}, c)
                tt.add({i: lin_name}, c * bias[j])
                tt.add({j: lin_name}, c * bias[i])
                tt.add({}, c * bias[i] * bias[j])

    quadratic(Cinv, Qb, Ec, "charge", "charge2")
    quadratic(Linv, Pbi, El, "flux", "flux2")

    def displacement(vec, phase_of, pref, up, down):
        for e in range(len(SS.edges)):
            amp = vec[e]
            if amp == 0:
                continue
            comp = _branch_composition(SS, e)
            ph = 2 * sy.pi * phase_of(e)
            c = -half * pref * amp
            cre = c * sy.cos(ph)
            cim = c * sy.sin(ph)
            plus = {nodes.index(node): (up if s > 0 else down) for node, s in comp}
            minus = {nodes.index(node): (up if s < 0 else down) for node, s in comp}
            tt.add(plus, cre, cim)       # e^{+i ph} * Pi^+
            tt.add(minus, cre, -cim)     # e^{-i ph} * Pi^-

    displacement(Jvec, lambda e: Pbm[e, e], Ej, "disp", "disp_adj")
    displacement(Pvec, lambda e: Qbt[e, 0], Ep, "pdisp", "pdisp_adj")

    # ---- node specs -----------------------------------------------------------------------
    node_specs = []
    for node in nodes:
        data = operator_data[node]
        spec = NodeSpec(node=node, basis=data["basis"], trunc=int(data["truncation"]),
                        cooper_disp=float(SS.cooper_disp[node]), fluxon_disp=float(SS.fluxon_disp[node]))
        if data["basis"] == "oscillator":
            spec.chg_osc = units.getPrefactor("ChgOsc")
            spec.flx_osc = units.getPrefactor("FlxOsc")
            if nodes.index(node) in regen:
                spec.regen = True
            else:
                spec.impedance_ohm = float(R(data["impedance"])) * units.getPrefactor("Impe")
        node_specs.append(spec)

    # ---- scalar outputs -------------------------------------------------------------------
    scalar_names, scalar_exprs = [], []

    def add_scalar(name, expr):
        if name not in scalar_names:
            scalar_names.append(name)
            scalar_exprs.append(sy.sympify(expr))

    for name, expr in (scalars or {}).items():
        if isinstance(expr, str):
            expr = SS.getSymbol(expr)
        add_scalar(name, R(expr))

    post = {}
    if resonator is not None:
        # inputs of getResonatorResponse (numerical_system.py:744-749)
        cpl = resonator["cpl_node"]
        idx = nodes.index(cpl)
        add_scalar("gC", R(SS.getSymbol("g%ir" % cpl)))
        add_scalar("frl", R(SS.getSymbol("f%irl" % cpl)))
        # a regenerated node carries the unit charge operator: <i|Q + q_b|j> / q(Z) = <i|Q~ + q_b/q(Z)|j>,
        # and only ratios of these matrix elements enter (numerical_system.py:755-762)
        if idx in regen:
            add_scalar("qb_cpl", Qb[idx] / (regen[idx]["q0"] if regen[idx]["stale"] else regen[idx]["q"]))
        else:
            add_scalar("qb_cpl", Qb[idx])
        post["resonator"] = {
            "cpl_node": int(cpl), "node_index": idx, "nmax": int(resonator.get("nmax", 100)),
            "op_factor": tt.fid(idx, "charge"),
            "gC": scalar_names.index("gC"), "frl": scalar_names.index("frl"),
            "qb": scalar_names.index("qb_cpl"),
        }

    # ---- auxiliary operators: Current / Voltage evaluables (numerical_system.py:596-694) -----------
    aux_keys, aux_coefs, aux_ops = [], [], []

    def aux_add(node_factor_names, re_expr, im_expr=0, fresh=False):
        if regen and any(ni in regen for ni in node_factor_names):
            expanded = list(_expand_regen(regen, node_factor_names, sy.sympify(re_expr), sy.sympify(im_expr), fresh))
        else:
            expanded = [(node_factor_names, re_expr, im_expr)]
        for names, re_e, im_e in expanded:
            key = [-1] * N
            for ni, name in names.items():
                key[ni] = tt.fid(ni, name)
            aux_keys.append(key)
            aux_coefs.append((sy.sympify(re_e), sy.sympify(im_e)))

    for spec in matrix_elements:
        begin = len(aux_keys)
        if spec["kind"] == "Voltage":
            node = spec.get("node")
            if node is None:
                raise Exception("No node specified for node voltage operator.")
            if node not in nodes:
                raise Exception("Node %s is not in the circuit." % repr(node))
            i = nodes.index(node)
            c = units.getPrefactor("Vop") * Cinv[i, i]                 # Vop * (Q_i + q_b,i) * C^-1_ii  (:678-681)
            aux_add({i: "charge"}, c)
            if Qb[i] != 0:
                aux_add({}, c * Qb[i])
            name = "getVoltageMatrixElement"
        elif spec["kind"] == "Current":
            edge = spec.get("edge")
            if isinstance(edge, list):
                edge = tuple(edge)
            if edge is None:
                raise Exception("No edge specified for branch current operator.")
            if edge not in SS.edges:
                raise Exception("Edge %s not in the circuit. Check that it is in the right direction." % repr(edge))
            e = SS.edges.index(edge)
            comp = _branch_composition(SS, e)
            if SS.CG.isInductiveEdge(edge):
                Pp = SS.Rnb * SS.Rinv * SS.node_vector
                Lb = R(SS.getInverseInductanceMatrix(mode='branch')[e, e])
                expr = Pp[e]
                args = expr.args if len(expr.atoms()) > 2 else [expr]
                for arg in args:                                         # IopL * sum coef * P_node * Linv_b[e,e]  (:608-620)
                    node = SS.node_map_rev[arg.args[1]]
                    aux_add({nodes.index(node): "flux"}, units.getPrefactor("IopL") * float(arg.args[0]) * Lb, fresh=True)
            elif SS.CG.isJosephsonEdge(edge):
                ph = 2 * sy.pi * Pbm[e, e]
                amp = half * units.getPrefactor("IopJ") * Jvec[e]      # 0.5j IopJ J (e^{+i ph} Pi+ - e^{-i ph} Pi-)  (:622-652)
                plus = {nodes.index(node): ("disp" if s_ > 0 else "disp_adj") for node, s_ in comp}
                minus = {nodes.index(node): ("disp" if s_ < 0 else "disp_adj") for node, s_ in comp}
                aux_add(plus, -amp * sy.sin(ph), amp * sy.cos(ph))
                aux_add(minus, -amp * sy.sin(ph), -amp * sy.cos(ph))
            else:
                raise Exception("Edge %s is not current-carrying" % repr(edge))
            name = "getCurrentMatrixElement"
        else:
            raise Exception("unknown matrix-element kind %r" % (spec["kind"],))
        if spec.get("elements") is None:
            raise Exception("No elements specified for %s matrix elements." % spec["kind"].lower())
        aux_ops.append({"name": name, "term_begin": begin, "term_end": len(aux_keys),
                        "elements": [[int(a), int(b)] for a, b in spec["elements"]]})

    # ---- compile --------------------------------------------------------------------------
    builder = ProgramBuilder(swept_syms)
    keys = [k for k in tt.order if not (tt.terms[k][0] == 0 and tt.terms[k][1] == 0)]
    program = builder.finish([tuple(tt.terms[k]) for k in keys], scalar_exprs, scalar_names, aux_coefs)
    return SweepPlan(
        nodes=node_specs,
        factors=list(tt.factors),
        terms=[list(k) for k in keys],
        program=program,
        swept_names=swept_names,
        sweep_specs=[dict(name=s["name"], start=float(s["start"]), end=float(s["end"]), N=int(s["N"]))
                     for s in sweep_specs],
        post=post,
        aux_terms=aux_keys,
        aux_ops=aux_ops,
        meta={"prefactors": {"Ec": Ec, "El": El, "Ej": Ej, "Ep": Ep}},
    )
