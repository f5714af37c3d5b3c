"""Flat coefficient program: sympy expressions -> SSA bytecode evaluated on the device.

The reference re-substitutes every symbolic matrix with ``sympy.Matrix.subs`` at every sweep
point (pyscqed/numerical_system.py:395-437) and re-propagates the parameterisation DAG
(pyscqed/parameters.py:1207-1240); that is 2-33 ms of pure Python per point.  Here every
sweep-dependent scalar is compiled ONCE into a straight-line program over the swept
parameters; kernel ``scqed_coef_eval`` (csrc/coef.cuh) runs it for all points.

Instruction format: int32 triples (op, a, b); instruction i writes register i (SSA).
Common sub-expressions are shared by hash-consing, so a register file of ``n_instr`` doubles
per point is the whole state.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import sympy as sy

# op codes -- keep in sync with csrc/coef.cuh (enum ScqedOp)
OP_CONST = 0   # r = consts[a]
OP_LOAD = 1    # r = params[a]
OP_ADD = 2
OP_SUB = 3
OP_MUL = 4
OP_DIV = 5
OP_NEG = 6
OP_SQRT = 7
OP_SIN = 8
OP_COS = 9
OP_EXP = 10
OP_LOG = 11
OP_POW = 12    # r = pow(r[a], r[b])
OP_POWI = 13   # r = r[a] ** b  (b is an immediate integer, may be negative)
OP_ABS = 14
OP_TAN = 15
OP_ATAN = 16
OP_TANH = 17
OP_SINH = 18
OP_COSH = 19

OP_NAMES = {v: k for k, v in dict(globals()).items() if k.startswith("OP_")}


@dataclass
class CoefProgram:
    """Compiled program.  ``coef_re/coef_im`` name the registers holding each complex term
    coefficient (-1 = exactly zero); ``scalar_regs`` name extra real per-point outputs."""
    n_swept: int
    instr: np.ndarray                      # int32 [n_instr, 3]
    consts: np.ndarray                     # float64 [n_consts]
    coef_re: np.ndarray                    # int32 [n_coef]
    coef_im: np.ndarray                    # int32 [n_coef]
    scalar_regs: np.ndarray                # int32 [n_scalar]
    scalar_names: list = field(default_factory=list)
    aux_re: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int32))   # auxiliary-operator terms
    aux_im: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int32))

    @property
    def n_instr(self) -> int:
        return int(self.instr.shape[0])

    @property
    def n_coef(self) -> int:
        return int(self.coef_re.shape[0])

    @property
    def n_scalar(self) -> int:
        return int(self.scalar_regs.shape[0])

    def to_dict(self) -> dict:
        return {
            "n_swept": int(self.n_swept),
            "instr": self.instr.astype(int).tolist(),
            # repr() round-trips doubles exactly
            "consts": [repr(float(c)) for c in self.consts],
            "coef_re": self.coef_re.astype(int).tolist(),
            "coef_im": self.coef_im.astype(int).tolist(),
            "scalar_regs": self.scalar_regs.astype(int).tolist(),
            "scalar_names": list(self.scalar_names),
            "aux_re": self.aux_re.astype(int).tolist(),
            "aux_im": self.aux_im.astype(int).tolist(),
        }

    @staticmethod
    def from_dict(d: dict) -> "CoefProgram":
        return CoefProgram(
            n_swept=int(d["n_swept"]),
            instr=np.asarray(d["instr"], dtype=np.int32).reshape(-1, 3),
            consts=np.asarray([float(c) for c in d["consts"]], dtype=np.float64),
            coef_re=np.asarray(d["coef_re"], dtype=np.int32),
            coef_im=np.asarray(d["coef_im"], dtype=np.int32),
            scalar_regs=np.asarray(d["scalar_regs"], dtype=np.int32),
            scalar_names=list(d.get("scalar_names", [])),
            aux_re=np.asarray(d.get("aux_re", []), dtype=np.int32),
            aux_im=np.asarray(d.get("aux_im", []), dtype=np.int32),
        )


class ProgramBuilder:
    """Hash-consing compiler from sympy expressions to the SSA bytecode."""

    def __init__(self, swept_symbols):
        self.swept = list(swept_symbols)
        self._sym_index = {s: i for i, s in enumerate(self.swept)}
        self.instr = []
        self.consts = []
        self._const_index = {}
        self._memo = {}       # (op, a, b) -> reg
        self._expr_memo = {}  # sympy expr -> reg

    # -- emission --------------------------------------------------------------
    def _emit(self, op, a=0, b=0):
        key = (op, a, b)
        if op in (OP_ADD, OP_MUL) and b < a:   # commutative: canonical order
            key = (op, b, a)
        r = self._memo.get(key)
        if r is None:
            r = len(self.instr)
            self.instr.append((key[0], key[1], key[2]))
            self._memo[key] = r
        return r

    def const(self, value) -> int:
        v = float(value)
        key = v.hex()
        ci = self._const_index.get(key)
        if ci is None:
            ci = len(self.consts)
            self.consts.append(v)
            self._const_index[key] = ci
        return self._emit(OP_CONST, ci, 0)

    # -- compilation -----------------------------------------------------------
    def compile(self, expr) -> int:
        """Return the register holding ``expr`` (a real-valued sympy expression or number)."""
        expr = sy.sympify(expr)
        r = self._expr_memo.get(expr)
        if r is None:
            r = self._compile(expr)
            self._expr_memo[expr] = r
        return r

    def _compile(self, e) -> int:
        if e.is_Number or e.is_NumberSymbol:
            if e.is_real is False:
                raise NotImplementedError("complex constant %r in a real coefficient expression" % (e,))
            return self.const(float(e))
        if e.is_Symbol:
            if e not in self._sym_index:
                raise KeyError("symbol %r is not a swept parameter; it should have been substituted" % (e,))
            return self._emit(OP_LOAD, self._sym_index[e], 0)
        if e.is_Add:
            # constant part last so that x + c and x - c share x
            pos, neg = [], []
            for arg in e.args:
                c, rest = arg.as_coeff_Mul()
                if c.is_negative and rest != 1:
                    neg.append(-arg)
                else:
                    pos.append(arg)
            if not pos:
                acc = self._emit(OP_NEG, self.compile(neg[0]), 0)
                neg = neg[1:]
            else:
                acc = self.compile(pos[0])
                for a in pos[1:]:
                    acc = self._emit(OP_ADD, acc, self.compile(a))
            for a in neg:
                acc = self._emit(OP_SUB, acc, self.compile(a))
            return acc
        if e.is_Mul:
            num, den = [], []
            for arg in e.args:
                if arg.is_Pow and arg.exp.is_Number and arg.exp.is_negative:
                    den.append(sy.Pow(arg.base, -arg.exp))
                else:
                    num.append(arg)
            if num:
                acc = self.compile(num[0])
                for a in num[1:]:
                    acc = self._emit(OP_MUL, acc, self.compile(a))
            else:
                acc = self.const(1.0)
            if den:
                dacc = self.compile(den[0])
                for a in den[1:]:
                    dacc = self._emit(OP_MUL, dacc, self.compile(a))
                acc = self._emit(OP_DIV, acc, dacc)
            return acc
        if e.is_Pow:
            base, ex = e.base, e.exp
            if ex.is_Integer:
                n = int(ex)
                if n == -1:
                    return self._emit(OP_DIV, self.const(1.0), self.compile(base))
                if n == 2:
                    rb = self.compile(base)
                    return self._emit(OP_MUL, rb, rb)
                return self._emit(OP_POWI, self.compile(base), n)
            if ex == sy.Rational(1, 2):
                return self._emit(OP_SQRT, self.compile(base), 0)
            if ex == sy.Rational(-1, 2):
                return self._emit(OP_DIV, self.const(1.0), self._emit(OP_SQRT, self.compile(base), 0))
            return self._emit(OP_POW, self.compile(base), self.compile(ex))
        unary = {
            sy.sin: OP_SIN, sy.cos: OP_COS, sy.exp: OP_EXP, sy.log: OP_LOG, sy.Abs: OP_ABS,
            sy.tan: OP_TAN, sy.atan: OP_ATAN, sy.tanh: OP_TANH, sy.sinh: OP_SINH, sy.cosh: OP_COSH,
        }
        if e.func in unary:
            if len(e.args) != 1:
                raise NotImplementedError("%r with %d arguments" % (e.func, len(e.args)))
            return self._emit(unary[e.func], self.compile(e.args[0]), 0)
        raise NotImplementedError("cannot lower sympy node %r (%s) to the coefficient program" % (e, type(e)))

    def finish(self, coef_pairs, scalar_exprs=(), scalar_names=(), aux_pairs=()) -> CoefProgram:
        """``coef_pairs`` / ``aux_pairs``: iterables of (re_expr, im_expr); ``scalar_exprs``: real expressions."""
        def regs(pairs):
            rr, ii = [], []
            for re_e, im_e in pairs:
                re_e = sy.sympify(re_e)
                im_e = sy.sympify(im_e)
                rr.append(-1 if re_e == 0 else self.compile(re_e))
                ii.append(-1 if im_e == 0 else self.compile(im_e))
            return rr, ii
        cre, cim = regs(coef_pairs)
        are, aim = regs(aux_pairs)
        sregs = [self.compile(s) for s in scalar_exprs]
        instr = np.asarray(self.instr, dtype=np.int32).reshape(-1, 3)
        return CoefProgram(
            n_swept=len(self.swept),
            instr=instr,
            consts=np.asarray(self.consts, dtype=np.float64),
            coef_re=np.asarray(cre, dtype=np.int32),
            coef_im=np.asarray(cim, dtype=np.int32),
            scalar_regs=np.asarray(sregs, dtype=np.int32),
            scalar_names=list(scalar_names),
            aux_re=np.asarray(are, dtype=np.int32),
            aux_im=np.asarray(aim, dtype=np.int32),
        )
