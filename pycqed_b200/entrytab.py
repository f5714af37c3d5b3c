"""Entry list for the dense assembly kernel (K2).

The sparsity pattern of  H(p) = sum_t c_t(p) kron_k F_{t,k}  does not depend on p.  It is expanded
once on the host (numpy Kronecker products in COO form) into

    entries   the distinct positions (i, j) that any term touches, sorted by i*n + j
    pairs     for every entry the (term, value) contributions:  H[i][j] = sum c_term * value

so that the kernel writes each structurally non-zero element exactly once (one thread per entry,
a handful of coefficient * value products) on top of a zero-filled matrix, instead of walking the
term and factor tables for all n^2 elements.  The reference forms the same sum with CSR algebra
(pyscqed/numerical_system.py:509-594); for the charge-basis circuits H has 0.2-3 % non-zeros.

Compile-time work on the plan (like the lowering); not used per sweep point on the host.
"""
from __future__ import annotations

import numpy as np

MAX_PAIRS = 40_000_000


def _factor_coo(M):
    i, j = np.nonzero(M)
    return i.astype(np.int64), j.astype(np.int64), M[i, j].astype(np.complex128)


def build_entry_table(plan, max_pairs: int = MAX_PAIRS):
    """None when the expansion would be too large, else dict(key int64 [E], ptr int32 [E+1],
    term int32 [NP], val complex128 [NP], dense bool)."""
    n = plan.hilbert_dim
    dims = plan.dims
    coo_cache = {}
    keys, terms, vals = [], [], []
    total = 0
    for t, fids in enumerate(plan.terms):
        rows = np.zeros(1, dtype=np.int64)
        cols = np.zeros(1, dtype=np.int64)
        v = np.ones(1, dtype=np.complex128)
        for k, fid in enumerate(fids):
            d = dims[k]
            if fid < 0:
                fi = fj = np.arange(d, dtype=np.int64)
                fv = np.ones(d, dtype=np.complex128)
            else:
                if fid not in coo_cache:
                    coo_cache[fid] = _factor_coo(plan.factor_matrix(fid))
                fi, fj, fv = coo_cache[fid]
            if rows.size * fi.size > max_pairs:
                return None
            rows = (rows[:, None] * d + fi[None, :]).reshape(-1)
            cols = (cols[:, None] * d + fj[None, :]).reshape(-1)
            v = (v[:, None] * fv[None, :]).reshape(-1)
        total += rows.size
        if total > max_pairs:
            return None
        keys.append(rows * n + cols)
        terms.append(np.full(rows.size, t, dtype=np.int32))
        vals.append(v)
    key = np.concatenate(keys)
    term = np.concatenate(terms)
    val = np.concatenate(vals)
    order = np.argsort(key, kind="stable")
    key, term, val = key[order], term[order], val[order]
    ukey, start = np.unique(key, return_index=True)
    ptr = np.concatenate([start, [key.size]]).astype(np.int64)
    if ptr[-1] >= 2 ** 31:
        return None
    return {"key": ukey.astype(np.int64), "ptr": ptr.astype(np.int32), "term": term, "val": val,
            "dense": bool(ukey.size * 4 >= n * n)}
