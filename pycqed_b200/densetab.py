"""Dense real term table for the fused small-n solver (S1).

For a Hilbert space that fits one SM's shared memory (n <= 224) the Kronecker term table
H(p) = sum_t c_t(p) kron_k F_{t,k}  is expanded ONCE on the host into a handful of real n x n
matrices G_g with real per-point coefficients:

    Re H(p) = sum_g alpha_g(p) G_g,      Im H(p) = sum_g beta_g(p) G_g

obtained from  c_t M_t = (a_t + i b_t)(R_t + i I_t):
  * every term whose coefficient does not depend on a swept parameter is folded into one constant
    matrix (the charging / inductive part of numerical_system.py:509-560),
  * numerically identical matrices (R_D = R_D^dagger, I_D^dagger = -I_D, ...) share one G_g and their
    coefficients are added by a few extra instructions appended to the coefficient program.

The S1 kernel then assembles H in shared memory with N_G coalesced 8-byte loads per element of the
lower triangle instead of walking the term / factor tables per element (which was a third of the
kernel).  For the fluxonium (C2) this is 3 real matrices instead of 6 complex term products.

This is compile-time work on the plan (like the lowering), not a host path of the sweep.
"""
from __future__ import annotations

import math

import numpy as np

from . import coefprog as cp
from .coefprog import CoefProgram

MAX_N = 224
_UNARY = {cp.OP_NEG: lambda x: -x, cp.OP_SQRT: math.sqrt, cp.OP_SIN: math.sin, cp.OP_COS: math.cos,
          cp.OP_EXP: math.exp, cp.OP_LOG: math.log, cp.OP_ABS: abs, cp.OP_TAN: math.tan, cp.OP_ATAN: math.atan,
          cp.OP_TANH: math.tanh, cp.OP_SINH: math.sinh, cp.OP_COSH: math.cosh}


def constant_registers(prog: CoefProgram):
    """(is_const[n_instr], value[n_instr]): registers that do not depend on a swept parameter."""
    n = prog.n_instr
    is_const = np.zeros(n, dtype=bool)
    val = np.zeros(n, dtype=np.float64)
    for i, (op, a, b) in enumerate(np.asarray(prog.instr).reshape(-1, 3).tolist()):
        try:
            if op == cp.OP_CONST:
                is_const[i], val[i] = True, prog.consts[a]
            elif op == cp.OP_LOAD:
                pass
            elif op in (cp.OP_ADD, cp.OP_SUB, cp.OP_MUL, cp.OP_DIV, cp.OP_POW):
                if is_const[a] and is_const[b]:
                    x, y = val[a], val[b]
                    val[i] = {cp.OP_ADD: x + y, cp.OP_SUB: x - y, cp.OP_MUL: x * y,
                              cp.OP_DIV: (x / y if y != 0 else math.inf), cp.OP_POW: x ** y}[op]
                    is_const[i] = True
            elif op == cp.OP_POWI:
                if is_const[a]:
                    val[i], is_const[i] = val[a] ** int(b), True
            elif op in _UNARY:
                if is_const[a]:
                    val[i], is_const[i] = _UNARY[op](val[a]), True
            else:
                raise ValueError("unknown opcode %d" % op)
        except (ValueError, OverflowError, ZeroDivisionError):
            is_const[i] = False          # leave it to the device
    return is_const, val


def _expand(plan, fids):
    out = np.ones((1, 1), dtype=np.complex128)
    for k, fid in enumerate(fids):
        m = np.eye(plan.dims[k], dtype=np.complex128) if fid < 0 else plan.factor_matrix(fid)
        out = np.kron(out, m)
    return out


def build_dense_table(plan, max_n: int = MAX_N):
    """Returns None when the plan is too large, else a dict with

    ``program``  the plan's coefficient program with the combination instructions appended,
    ``table``    float64: the ``n_full`` general matrices, element (i, j) of G_g at ``g*n*n + j*n + i`` (column-major),
                 followed by the ``n_diag`` purely diagonal ones as n-vectors,
    ``re_regs`` / ``im_regs``  int32 [n_g] registers of alpha_g / beta_g (-1 = identically zero),
    ``gmax``     float64 [n_g] max |G_g| (the kernel's realness test: sum |beta_g| gmax_g < 1e-12).
    """
    n = plan.hilbert_dim
    if n > max_n or not plan.terms:
        return None
    prog = plan.program
    is_const, val = constant_registers(prog)
    instr = np.asarray(prog.instr, dtype=np.int32).reshape(-1, 3).tolist()
    consts = list(np.asarray(prog.consts, dtype=np.float64))

    const_re = np.zeros((n, n), dtype=np.float64)
    const_im = np.zeros((n, n), dtype=np.float64)
    mats = []                                  # [matrix, [(sign, reg)] for Re H, [(sign, reg)] for Im H]

    def slot(M):
        scale = np.abs(M).max()
        for e in mats:
            if e[0].shape == M.shape:
                if np.abs(e[0] - M).max() <= 1e-13 * scale:
                    return e, 1
                if np.abs(e[0] + M).max() <= 1e-13 * scale:
                    return e, -1
        e = [M, [], []]
        mats.append(e)
        return e, 1

    def contribute(M, reg, sign, part):
        """sign * r[reg] * M goes to the real (part = 0) or imaginary (part = 1) part of H."""
        if reg < 0 or not np.any(M):
            return
        if is_const[reg]:
            (const_re if part == 0 else const_im)[...] += sign * val[reg] * M
            return
        e, s = slot(M)
        e[1 + part].append((sign * s, reg))

    for t, fids in enumerate(plan.terms):
        M = _expand(plan, fids)
        R, I = np.ascontiguousarray(M.real), np.ascontiguousarray(M.imag)
        a, b = int(prog.coef_re[t]), int(prog.coef_im[t])
        contribute(R, a, +1, 0)      # Re += a R - b I
        contribute(I, b, -1, 0)
        contribute(I, a, +1, 1)      # Im += a I + b R
        contribute(R, b, +1, 1)

    def emit(op, a=0, b=0):
        instr.append([op, a, b])
        return len(instr) - 1

    def combine(pairs):
        # cancel +r / -r, then a short add/sub chain
        net = {}
        for s, r in pairs:
            net[r] = net.get(r, 0) + s
        acc = -1
        for r, s in net.items():
            if s == 0:
                continue
            term = r
            if abs(s) != 1:
                consts.append(float(abs(s)))
                term = emit(cp.OP_MUL, emit(cp.OP_CONST, len(consts) - 1), r)
            if acc < 0:
                acc = term if s > 0 else emit(cp.OP_NEG, term)
            else:
                acc = emit(cp.OP_ADD if s > 0 else cp.OP_SUB, acc, term)
        return acc

    table, re_regs, im_regs = [], [], []
    one = -1
    for M, part in ((const_re, 0), (const_im, 1)):
        if np.any(M):
            if one < 0:
                consts.append(1.0)
                one = emit(cp.OP_CONST, len(consts) - 1)
            table.append(M)
            re_regs.append(one if part == 0 else -1)
            im_regs.append(one if part == 1 else -1)
    for M, re_pairs, im_pairs in mats:
        rr, ii = combine(re_pairs), combine(im_pairs)
        if rr < 0 and ii < 0:
            continue
        table.append(M)
        re_regs.append(rr)
        im_regs.append(ii)
    if not table:
        return None
    # purely diagonal matrices (charge-basis Q, Q^2; the projector terms of a swept-impedance oscillator)
    # go last and are stored as n-vectors: they only touch the n diagonal elements
    is_diag = [not np.any(M - np.diag(np.diag(M))) for M in table]
    order = [i for i, dg in enumerate(is_diag) if not dg] + [i for i, dg in enumerate(is_diag) if dg]
    table = [table[i] for i in order]
    re_regs = [re_regs[i] for i in order]
    im_regs = [im_regs[i] for i in order]
    n_diag = int(sum(is_diag))
    n_full = len(table) - n_diag
    tab = np.concatenate([np.ascontiguousarray(M.T).reshape(-1) for M in table[:n_full]] +
                         [np.ascontiguousarray(np.diag(M)).reshape(-1) for M in table[n_full:]] + [np.zeros(0)])
    program = CoefProgram(
        n_swept=prog.n_swept, instr=np.asarray(instr, dtype=np.int32).reshape(-1, 3),
        consts=np.asarray(consts, dtype=np.float64), coef_re=prog.coef_re, coef_im=prog.coef_im,
        scalar_regs=prog.scalar_regs, scalar_names=list(prog.scalar_names), aux_re=prog.aux_re, aux_im=prog.aux_im)
    ii, jj = np.indices((n, n))
    tridiagonal = all(not np.any(M[np.abs(ii - jj) > 1]) for M in table[:n_full])
    return {"program": program, "table": np.ascontiguousarray(tab, dtype=np.float64), "n_full": n_full, "n_diag": n_diag,
            "tridiagonal": bool(tridiagonal and n >= 3),
            "re_regs": np.asarray(re_regs, dtype=np.int32), "im_regs": np.asarray(im_regs, dtype=np.int32),
            "gmax": np.asarray([np.abs(M).max() for M in table], dtype=np.float64)}
