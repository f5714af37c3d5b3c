"""CPU oracle for the sweep hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module; ``pycqed_b200/`` never does (the product path has no CPU fallback).

It restates, with numpy/scipy, what the reference does at every sweep point:

  coefficients   pyscqed/numerical_system.py:395-437 (``_postsub``) -- here by interpreting the
                 plan's coefficient program on the host
  assembly       pyscqed/numerical_system.py:509-594 (``getHamiltonian``): a sum of scaled
                 Kronecker products in CSR sparse algebra (what QuTiP's ``tensor``/``+``/``*``
                 do, pyscqed/numerical_system.py:226-251), then ``tidyup(1e-12)`` (:594)
  eigensolve     pyscqed/util.py:153 ``scipy.linalg.eigh(H, subset_by_index=(0, k-1))``
                 (LAPACK zheevr), sorted / normalised as util.py:156-176;
                 pyscqed/util.py:102 ``scipy.sparse.linalg.eigsh`` for the sparse variant
  resonator      pyscqed/numerical_system.py:736-784 (``getResonatorResponse``)
  result layout  pyscqed/parameters.py:1060-1068,1146-1157 (``getSweepResult`` transposition)

Third-party arithmetic the reference delegates to and that is not under /root/reference:
QuTiP 5.0.3.post1 CSR algebra (restated above), SciPy -> LAPACK ``zheevr`` / ARPACK (called
directly, same entry points as the reference).

Parity pinning: the reference has NO tests or golden vectors for this path,
so the oracle is pinned against outputs of the reference itself, run unmodified in the build
container over ``oracle/shims`` -- see ``tests/golden/make_golden.py`` and
``tests/test_oracle_golden.py`` / ``tests/test_dropin_vs_reference.py``.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import scipy.sparse as sp
import scipy.sparse.linalg

QOBJ_ATOL = 1e-12   # numerical_system.py:19


# ------------------------------------------------------------------------------------------------
# coefficient program
# ------------------------------------------------------------------------------------------------
def _run_program(program, params):
    """Interpret the SSA bytecode for all points at once -> register file [n_instr, n_pts]."""
    params = np.asarray(params, dtype=np.float64)
    if params.ndim != 2 or params.shape[1] != program.n_swept:
        raise ValueError("params must be [n_pts, %d], got %r" % (program.n_swept, params.shape))
    n_pts = params.shape[0]
    regs = np.zeros((program.n_instr, n_pts), dtype=np.float64)
    unary = {6: np.negative, 7: np.sqrt, 8: np.sin, 9: np.cos, 10: np.exp, 11: np.log, 14: np.abs, 15: np.tan,
             16: np.arctan, 17: np.tanh, 18: np.sinh, 19: np.cosh}
    binary = {2: np.add, 3: np.subtract, 4: np.multiply, 5: np.divide, 12: np.power}
    with np.errstate(all="ignore"):
        for i, (op, a, b) in enumerate(program.instr.tolist()):
            if op == 0:
                regs[i] = program.consts[a]
            elif op == 1:
                regs[i] = params[:, a]
            elif op in binary:
                regs[i] = binary[op](regs[a], regs[b])
            elif op in unary:
                regs[i] = unary[op](regs[a])
            elif op == 13:
                regs[i] = np.power(regs[a], float(b))
            else:
                raise ValueError("unknown opcode %d" % op)
    return regs


def _complex_outputs(regs, re_ids, im_ids):
    out = np.zeros((regs.shape[1], len(re_ids)), dtype=np.complex128)
    for t in range(len(re_ids)):
        re = regs[re_ids[t]] if re_ids[t] >= 0 else 0.0
        im = regs[im_ids[t]] if im_ids[t] >= 0 else 0.0
        out[:, t] = re + 1j * im
    return out


def eval_program(program, params):
    """:param params: [n_pts, n_swept] float64
    :return: (coef complex128 [n_pts, n_coef], scalars float64 [n_pts, n_scalar])
    """
    regs = _run_program(program, params)
    coef = _complex_outputs(regs, program.coef_re, program.coef_im)
    scal = np.zeros((regs.shape[1], program.n_scalar), dtype=np.float64)
    for s in range(program.n_scalar):
        scal[:, s] = regs[program.scalar_regs[s]]
    return coef, scal


def eval_aux_coefficients(program, params):
    """Coefficients of the auxiliary-operator terms (Current / Voltage evaluables) [n_pts, n_aux]."""
    return _complex_outputs(_run_program(program, params), program.aux_re, program.aux_im)


# ------------------------------------------------------------------------------------------------
# assembly
# ------------------------------------------------------------------------------------------------
def _tidyup(m, atol=QOBJ_ATOL):
    m = sp.csr_matrix(m, dtype=np.complex128)
    re = m.data.real.copy()
    im = m.data.imag.copy()
    re[np.abs(re) < atol] = 0.0
    im[np.abs(im) < atol] = 0.0
    m.data = re + 1j * im
    m.eliminate_zeros()
    return m


class PlanOracle:
    """Holds the Kronecker-expanded term operators (the reference's ``circ_operators`` products)
    so that a per-point assembly is the same CSR scale-and-add the reference performs."""

    def __init__(self, plan):
        self.plan = plan
        self.dims = plan.dims
        self.n = plan.hilbert_dim
        self._factors = [sp.csr_matrix(plan.factor_matrix(f)) for f in range(len(plan.factors))]
        self._eyes = [sp.identity(d, dtype=np.complex128, format="csr") for d in self.dims]
        self.term_ops = [self._expand(t) for t in plan.terms]
        self.aux_term_ops = [self._expand(t) for t in getattr(plan, "aux_terms", [])]
        # nodes lowered in a rotated basis (swept oscillator impedance): T maps plan-basis vectors to
        # the reference's Fock basis, v = T v'
        self.T = None
        xf = plan.node_transforms() if hasattr(plan, "node_transforms") else {}
        if xf:
            T = np.ones((1, 1))
            for k, d in enumerate(self.dims):
                T = np.kron(T, xf[k] if k in xf else np.eye(d))
            self.T = T

    def _expand(self, fids):
        out = None
        for k, fid in enumerate(fids):
            m = self._eyes[k] if fid < 0 else self._factors[fid]
            out = m if out is None else sp.kron(out, m, format="csr")
        return sp.csr_matrix(out)

    def expand_factor(self, fid):
        ni = self.plan.factors[fid][0]
        fids = [-1] * len(self.dims)
        fids[ni] = fid
        return self._expand(fids)

    def coefficients(self, params):
        return eval_program(self.plan.program, np.atleast_2d(np.asarray(params, dtype=np.float64)))

    def hamiltonian(self, coef_row, tidy=True):
        H = None
        for c, op in zip(coef_row, self.term_ops):
            H = c * op if H is None else H + c * op
        H = sp.csr_matrix(H)
        return _tidyup(H) if tidy else H

    def diag_dense(self, H, k, get_vectors=False):
        """util.py:127-176."""
        ret = scipy.linalg.eigh(H.toarray(), eigvals_only=(not get_vectors), subset_by_index=(0, k - 1))
        if not get_vectors:
            ret.sort()
            return ret
        E, V = ret
        order = np.argsort(E, kind="stable")
        E = E[order]
        V = V[:, order]
        V = V / np.linalg.norm(V, axis=0, keepdims=True)
        return E, V

    def diag_sparse(self, H, k, get_vectors=False, **opts):
        """util.py:76-125 (caller supplies e.g. sigma=-300, tol=0)."""
        ret = scipy.sparse.linalg.eigsh(sp.csr_matrix(H), k=k, return_eigenvectors=get_vectors, **opts)
        if not get_vectors:
            ret.sort()
            return ret
        E, V = ret
        order = np.argsort(E, kind="stable")
        return E[order], V[:, order] / np.linalg.norm(V[:, order], axis=0, keepdims=True)

    # -- Current / Voltage operators and their matrix elements ----------------------------------
    def aux_operator(self, aux_coef_row, op_index):
        """getCurrentOperator / getVoltageOperator (numerical_system.py:596-654,669-681) as CSR."""
        op = self.plan.aux_ops[op_index]
        M = sp.csr_matrix((self.n, self.n), dtype=np.complex128)
        for t in range(op["term_begin"], op["term_end"]):
            M = M + aux_coef_row[t] * self.aux_term_ops[t]
        return sp.csr_matrix(M)

    def matrix_elements(self, params, V):
        """<V_i|Op|V_j> (complex) for every requested pair, op after op, as ``plan.aux_pairs()`` orders
        them; ``V`` is [n_pts, n, k].  The reference's evaluables keep ``.real``
        (numerical_system.py:656-667,683-694)."""
        params = np.atleast_2d(np.asarray(params, dtype=np.float64))
        ac = eval_aux_coefficients(self.plan.program, params)
        if self.T is not None:                                   # Fock basis in, plan basis inside
            V = np.einsum("ba,pbk->pak", self.T, np.asarray(V))
        out = []
        for p in range(params.shape[0]):
            row = []
            for oi, op in enumerate(self.plan.aux_ops):
                M = self.aux_operator(ac[p], oi)
                MV = M @ V[p]
                for i, j in op["elements"]:
                    row.append(np.vdot(V[p][:, i], MV[:, j]))
            out.append(row)
        return np.asarray(out, dtype=np.complex128)

    # -- resonator ----------------------------------------------------------------------------
    def resonator_response(self, E, V, scal_row):
        """numerical_system.py:736-784, literally (bare-order bookkeeping included)."""
        spec = self.plan.post["resonator"]
        gC = scal_row[spec["gC"]]
        wrl = scal_row[spec["frl"]]
        qb = scal_row[spec["qb"]]
        Op = self.expand_factor(spec["op_factor"]) + qb * sp.identity(self.n, dtype=np.complex128, format="csr")
        return resonator_response(np.asarray(E), V, Op, gC, wrl, spec["nmax"])

    # -- whole sweep -----------------------------------------------------------------------------
    def sweep(self, params, k, get_vectors=False, resonator=False, sparse_opts=None):
        params = np.atleast_2d(np.asarray(params, dtype=np.float64))
        coef, scal = eval_program(self.plan.program, params)
        n_pts = params.shape[0]
        Es = np.zeros((n_pts, k))
        Vs = np.zeros((n_pts, self.n, k), dtype=np.complex128) if get_vectors or resonator else None
        Rs = []
        for i in range(n_pts):
            H = self.hamiltonian(coef[i])
            if sparse_opts is not None:
                out = self.diag_sparse(H, k, get_vectors or resonator, **sparse_opts)
            else:
                out = self.diag_dense(H, k, get_vectors or resonator)
            if get_vectors or resonator:
                Es[i], Vs[i] = out
                if resonator:
                    Rs.append(self.resonator_response(Es[i], Vs[i], scal[i]))
                if self.T is not None:
                    Vs[i] = self.T @ Vs[i]
            else:
                Es[i] = out
        return {"E": Es, "V": Vs, "Erwa": (np.array(Rs) if resonator else None), "scalars": scal}


def pauli_coefficients(E0, E1, Op):
    """util.pauliCoefficients (util.py:300-391) in float64 (the reference computes in complex64):
    local-basis reduction of the two lowest levels.  ``Op`` [n_pts, 2, 2] is the computational-basis
    operator in the {|0>, |1>} eigenbasis.  Returns hx, hy, hz."""
    ox = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    oy = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
    oz = np.array([[1, 0], [0, -1]], dtype=np.complex128)
    hx, hy, hz = [], [], []
    for i in range(len(E0)):
        _, Vl = np.linalg.eigh(np.asarray(Op[i], dtype=np.complex128))
        Vl = np.array([np.abs(Vl[0, k]) / Vl[0, k] * Vl[:, k] for k in (0, 1)])
        U = Vl.T
        Hq = U.conj().T @ np.diag([E0[i], E1[i]]).astype(np.complex128) @ U
        hx.append(np.real(0.5 * np.trace(Hq @ ox)))
        hy.append(np.real(0.5 * np.trace(Hq @ oy)))
        hz.append(np.real(0.5 * np.trace(Hq @ oz)))
    return np.array(hx), np.array(hy), np.array(hz)


def resonator_response(E, V, Op, gC, wrl, nmax=100):
    """RWA strips (numerical_system.py:751-784).  ``V`` is [n, k] (columns = kets)."""
    E = E - E[0]
    tmax = len(E)
    nmax = nmax + 1 + tmax
    OV = Op @ V

    def mel(i, j):
        return np.vdot(V[:, i], OV[:, j])

    norm = mel(0, 1)
    g_list = np.array([mel(i, i + 1) for i in range(tmax - 1)])
    g_list = gC * np.abs(g_list / norm)
    order = np.zeros(tmax, dtype=int)
    diag_bare = np.array([-i * wrl + E[i] for i in range(tmax)])
    eigensolver_order = np.linalg.eigvalsh(np.diag(diag_bare))
    for i in range(tmax):
        index, = np.where(eigensolver_order == diag_bare[i])
        order[i] = index[0]
    Erwa = [0.0] * nmax
    for n in range(nmax):
        d = np.array([(n - i) * wrl + E[i] for i in range(tmax)])
        o = np.array([g_list[i] * np.sqrt((n - i) * (n - i > 0)) for i in range(tmax - 1)])
        e = scipy.linalg.eigvalsh(np.diag(d) + np.diag(o, 1) + np.diag(o, -1))
        Erwa[n] = np.array([e[i].real for i in order])
    return np.array(Erwa)


# ------------------------------------------------------------------------------------------------
# comparison helpers used by the parity tests
# ------------------------------------------------------------------------------------------------
def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def eig_rel_err(E, Eref, floor_frac=1e-3):
    """Relative eigenvalue error with a per-point floor: |dE_i| / max(|Eref_i|, floor_frac * max_j |Eref_j|).
    A level that is exactly 0 (e.g. the Averin coupler at integer gate charge with the junction switched
    off) has no meaningful relative error; every eigensolver, LAPACK included, is accurate to eps*||H||."""
    E = np.atleast_2d(np.asarray(E, dtype=np.float64))
    Eref = np.atleast_2d(np.asarray(Eref, dtype=np.float64))
    floor = floor_frac * np.max(np.abs(Eref), axis=1, keepdims=True)
    return float(np.max(np.abs(E - Eref) / np.maximum(np.abs(Eref), np.maximum(floor, 1e-300))))


def cluster_indices(E, rtol=1e-7, atol=1e-9):
    """Group (near-)degenerate levels: consecutive values closer than atol + rtol*|E|."""
    groups, cur = [], [0]
    for i in range(1, len(E)):
        if abs(E[i] - E[i - 1]) <= atol + rtol * max(abs(E[i]), abs(E[i - 1])):
            cur.append(i)
        else:
            groups.append(cur)
            cur = [i]
    groups.append(cur)
    return groups


def subspace_overlap(Va, Vb, E, rtol=1e-7, atol=1e-9, E_next=None):
    """min over degenerate clusters of the smallest singular value of Va_c^H Vb_c
    (== |<a|b>| for a non-degenerate level; invariant under phase and rotation inside a
    degenerate subspace).

    The cluster containing level k-1 is compared like every other one UNLESS the k-truncation cuts it:
    ``E_next`` (the reference's level k, i.e. the first one not returned) tells.  The floating CSFQ has
    exact doublets; when the partner of level k-1 is level k any vector of that eigenspace is a correct
    answer and only the residual can be checked.  Without ``E_next`` (k == n, nothing is cut) every
    cluster is compared; for k < n ``E_next`` is required."""
    worst = 1.0
    k = Va.shape[1]
    E = np.asarray(E, dtype=np.float64)
    if k < Va.shape[0] and E_next is None:
        raise ValueError("subspace_overlap: pass E_next (the reference's level k) when k < n")
    for g in cluster_indices(E, rtol, atol):
        if g[-1] == k - 1 and k < Va.shape[0]:
            nxt = float(np.atleast_1d(E_next)[0])
            if abs(nxt - E[k - 1]) <= atol + rtol * max(abs(nxt), abs(E[k - 1])):
                continue                    # the truncation cuts a degenerate cluster
        A, B = Va[:, g], Vb[:, g]
        if len(g) > 1:
            # SPANS are compared: ARPACK (the reference's sparse path, util.py:102) returns vectors of an exactly
            # degenerate cluster that are not mutually orthogonal (measured: <v2|v3> = 3e-3 on the floating CSFQ
            # doublets at tol=0), so each side's cluster is orthonormalised first.  Orthonormality of the vectors
            # under test is asserted separately by the callers.
            A, B = np.linalg.qr(A)[0], np.linalg.qr(B)[0]
        S = np.linalg.svd(A.conj().T @ B, compute_uv=False)
        worst = min(worst, float(S.min()))
    return worst


def residual_norms(H, E, V):
    """max-abs residual of each column of V: ||H v - E v||_inf (H dense or sparse)."""
    R = H @ V - V * np.asarray(E)[None, :]
    return np.max(np.abs(R), axis=0)
