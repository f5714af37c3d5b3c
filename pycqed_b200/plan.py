"""SweepPlan: the device-side description of H(p) = sum_t c_t(p) * kron_k O_{t,k}.

Three tables replace the reference's full-space QuTiP operators
(pyscqed/numerical_system.py:185-253) and per-point sympy substitution (:395-437):

  factor table   small d_k x d_k complex matrices, one set per circuit node
  term table     per Hamiltonian term one factor id per node (-1 = identity)
  coefficient    straight-line program giving every complex c_t and a few real
  program        scalars (g_C, f_rl, charging energies ...) from the swept values

A plan is pure data; it serialises to a small JSON document (factors are stored as
*specifications* and rebuilt with :mod:`pycqed_b200.operators`), which is how the GPU tests
and ``bench.py`` obtain the reference circuits on a box without the reference installed.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field

import numpy as np

from . import operators as ops
from .coefprog import CoefProgram

#: factor names; the first six are the reference's ``circ_operators[node]`` keys
#: (numerical_system.py:226-251), the last two are same-node products Q.Q and P.P.
FACTOR_NAMES = ("charge", "flux", "disp", "disp_adj", "pdisp", "pdisp_adj", "charge2", "flux2")


@dataclass
class NodeSpec:
    node: int                 # circuit node label (SS.nodes entry)
    basis: str                # 'charge' | 'oscillator' | 'flux'
    trunc: int
    impedance_ohm: float = 0.0   # oscillator only: Zosc<node> * Impe
    chg_osc: float = 0.0         # Units.getPrefactor('ChgOsc')
    flx_osc: float = 0.0         # Units.getPrefactor('FlxOsc')
    cooper_disp: float = 1.0
    fluxon_disp: float = 1.0
    regen: bool = False          # oscillator whose impedance is swept: unit operators in the eigenbasis of a+a^dag

    @property
    def dim(self) -> int:
        return ops.node_dimension(self.basis, self.trunc)

    def transform(self):
        """Real orthogonal W (columns = eigenvectors of a + a^dag) that rotates this node's coordinates
        back to the Fock basis, or None (see operators.oscillator_regen_set)."""
        if not self.regen:
            return None
        return ops.oscillator_regen_set(self.trunc)["W"]

    def operator_set(self) -> dict:
        if self.regen:
            if self.basis != "oscillator":
                raise Exception("only oscillator nodes are regenerated")
            rs = ops.oscillator_regen_set(self.trunc)
            return {k: rs[k] for k in ("charge", "flux", "charge2", "flux2")}
        if self.basis == "charge":
            t = ops.charge_basis(self.trunc)
        elif self.basis == "oscillator":
            t = ops.oscillator_basis(self.trunc, self.impedance_ohm, self.chg_osc, self.flx_osc,
                                     self.cooper_disp, self.fluxon_disp)
        elif self.basis == "flux":
            t = ops.flux_basis(self.trunc)
        else:
            raise Exception("Unrecognized basis representation '%s'." % repr(self.basis))
        out = dict(zip(ops.FACTOR_KINDS, t))
        out["charge2"] = out["charge"] @ out["charge"]
        out["flux2"] = out["flux"] @ out["flux"]
        return out


@dataclass
class SweepPlan:
    nodes: list                       # [NodeSpec]
    factors: list                     # [(node_index, name)]
    terms: list                       # [[factor id per node]]; coefficient index == term index
    program: CoefProgram
    swept_names: list
    sweep_specs: list = field(default_factory=list)   # [{'name','start','end','N'}] in addSweep order
    post: dict = field(default_factory=dict)          # e.g. {'resonator': {...}}
    aux_terms: list = field(default_factory=list)     # extra operator terms [[factor id per node]]
    aux_ops: list = field(default_factory=list)       # [{'name', 'term_begin', 'term_end', 'elements': [[i, j]]}]
    meta: dict = field(default_factory=dict)
    _opsets: dict = field(default_factory=dict, repr=False)

    # -- geometry ------------------------------------------------------------------------
    @property
    def dims(self):
        return [n.dim for n in self.nodes]

    @property
    def hilbert_dim(self) -> int:
        return int(np.prod(self.dims)) if self.nodes else 1

    @property
    def n_terms(self) -> int:
        return len(self.terms)

    def factor_matrix(self, fid: int) -> np.ndarray:
        ni, name = self.factors[fid]
        if name.startswith("proj:"):                      # |m><m| of a regenerated oscillator node
            d = self.nodes[ni].dim
            m = np.zeros((d, d), dtype=np.complex128)
            m[int(name[5:]), int(name[5:])] = 1.0
            return m
        if ni not in self._opsets:
            self._opsets[ni] = self.nodes[ni].operator_set()
        return self._opsets[ni][name]

    def node_transforms(self) -> dict:
        """{node index: W} for the nodes solved in a rotated basis (eigenvectors are rotated back)."""
        return {i: nd.transform() for i, nd in enumerate(self.nodes) if nd.regen}

    def factor_kind(self, fid: int) -> str:
        return ops.classify(self.factor_matrix(fid))

    def aux_pairs(self) -> list:
        """Flat list [term_begin, term_end, i, j] of all requested matrix elements, op after op."""
        out = []
        for op in self.aux_ops:
            for i, j in op["elements"]:
                out.append([int(op["term_begin"]), int(op["term_end"]), int(i), int(j)])
        return out

    def aux_as_terms(self, op_index: int) -> "SweepPlan":
        """A plan whose term table is auxiliary operator ``op_index`` (so K1 + K2 assemble
        getCurrentOperator / getVoltageOperator, numerical_system.py:596-654,669-681, like an H)."""
        op = self.aux_ops[op_index]
        tb, te = int(op["term_begin"]), int(op["term_end"])
        prog = self.program
        sub = CoefProgram(n_swept=prog.n_swept, instr=prog.instr, consts=prog.consts,
                          coef_re=prog.aux_re[tb:te].copy(), coef_im=prog.aux_im[tb:te].copy(),
                          scalar_regs=prog.scalar_regs, scalar_names=list(prog.scalar_names))
        return SweepPlan(nodes=self.nodes, factors=self.factors, terms=[list(t) for t in self.aux_terms[tb:te]],
                         program=sub, swept_names=self.swept_names, sweep_specs=self.sweep_specs, meta=dict(self.meta))

    def sweep_grid(self) -> np.ndarray:
        """[n_pts, n_swept] row-major, C-order over the addSweep order (last = fastest),
        exactly the reference's collapsed grid (pyscqed/parameters.py:861-886)."""
        if not self.sweep_specs:
            return np.zeros((1, 0), dtype=np.float64)
        axes = [np.linspace(s["start"], s["end"], int(s["N"])) for s in self.sweep_specs]
        mesh = np.meshgrid(*axes, indexing="ij")
        by_name = {s["name"]: m.reshape(-1) for s, m in zip(self.sweep_specs, mesh)}
        return np.ascontiguousarray(np.stack([by_name[n] for n in self.swept_names], axis=1))

    # -- structure queries used to pick a solver -------------------------------------------
    def is_tridiagonal(self) -> bool:
        """True when H is Hermitian tridiagonal for every p (single node, all factors band1)."""
        if len(self.nodes) != 1:
            return False
        for t in self.terms:
            if t[0] >= 0 and self.factor_kind(t[0]) == "dense":
                return False
        return True

    # -- (de)serialisation ---------------------------------------------------------------
    def to_dict(self) -> dict:
        return {
            "format": "pycqed_b200.SweepPlan/1",
            "nodes": [
                {"node": int(n.node), "basis": n.basis, "trunc": int(n.trunc),
                 "impedance_ohm": repr(float(n.impedance_ohm)), "chg_osc": repr(float(n.chg_osc)),
                 "flx_osc": repr(float(n.flx_osc)), "cooper_disp": repr(float(n.cooper_disp)),
                 "fluxon_disp": repr(float(n.fluxon_disp)), "regen": bool(n.regen)}
                for n in self.nodes
            ],
            "factors": [[int(a), str(b)] for a, b in self.factors],
            "terms": [[int(x) for x in t] for t in self.terms],
            "program": self.program.to_dict(),
            "swept_names": list(self.swept_names),
            "sweep_specs": [
                {"name": s["name"], "start": repr(float(s["start"])), "end": repr(float(s["end"])), "N": int(s["N"])}
                for s in self.sweep_specs
            ],
            "post": self.post,
            "aux_terms": [[int(x) for x in t] for t in self.aux_terms],
            "aux_ops": self.aux_ops,
            "meta": self.meta,
        }

    @staticmethod
    def from_dict(d: dict) -> "SweepPlan":
        nodes = [
            NodeSpec(node=int(n["node"]), basis=n["basis"], trunc=int(n["trunc"]),
                     impedance_ohm=float(n["impedance_ohm"]), chg_osc=float(n["chg_osc"]),
                     flx_osc=float(n["flx_osc"]), cooper_disp=float(n["cooper_disp"]),
                     fluxon_disp=float(n["fluxon_disp"]), regen=bool(n.get("regen", False)))
            for n in d["nodes"]
        ]
        return SweepPlan(
            nodes=nodes,
            factors=[(int(a), str(b)) for a, b in d["factors"]],
            terms=[[int(x) for x in t] for t in d["terms"]],
            program=CoefProgram.from_dict(d["program"]),
            swept_names=list(d["swept_names"]),
            sweep_specs=[
                {"name": s["name"], "start": float(s["start"]), "end": float(s["end"]), "N": int(s["N"])}
                for s in d.get("sweep_specs", [])
            ],
            post=d.get("post", {}),
            aux_terms=[[int(x) for x in t] for t in d.get("aux_terms", [])],
            aux_ops=d.get("aux_ops", []),
            meta=d.get("meta", {}),
        )

    def save(self, path: str) -> None:
        with open(path, "w") as fh:
            json.dump(self.to_dict(), fh, indent=1)

    @staticmethod
    def load(path: str) -> "SweepPlan":
        with open(path) as fh:
            return SweepPlan.from_dict(json.load(fh))

    def with_sweep(self, specs) -> "SweepPlan":
        """Same circuit, different grid (the swept names must match)."""
        d = self.to_dict()
        d["sweep_specs"] = [
            {"name": s["name"], "start": repr(float(s["start"])), "end": repr(float(s["end"])), "N": int(s["N"])}
            for s in specs
        ]
        return SweepPlan.from_dict(d)
