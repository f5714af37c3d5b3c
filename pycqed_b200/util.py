"""Consumers of the sweep results, same names and argument meaning as ``pyscqed.util``.

Only the helpers that sit directly on the hot path's outputs are here:

  getResonatorShift / getCircuitLambShift / getACStarkShift   util.py:178-240 (index arithmetic on
                                                              the ``Resonator`` strips, vectorised)
  createSubspaceOperators, pauliCoefficients                  util.py:286-391 (local-basis reduction;
                                                              the per-point loop runs as one kernel)
  getEigenValuesAndVectors                                    util.py:393-402
"""
from __future__ import annotations

import numpy as np

from . import engine as _engine


def getResonatorShift(Erwa):
    """``ret[q, n]``: resonator frequency with the circuit in state q at n photons (util.py:218-240).
    ``Erwa`` is indexed [level, strip] as ``getSweep(..., evaluable='Resonator')`` returns it per point."""
    Erwa = np.asarray(Erwa)
    levels = Erwa.shape[0]
    photons = Erwa.shape[1] - levels
    q = np.arange(levels - 1)[:, None]
    n = np.arange(photons)[None, :]
    return Erwa[q, n + q + 1] - Erwa[q, n + q]


def getCircuitLambShift(Erwa):
    """``ret[q]``: dressed level q at zero photons (util.py:202-216)."""
    Erwa = np.asarray(Erwa)
    i = np.arange(Erwa.shape[0])
    return Erwa[i, i]


def getACStarkShift(Erwa):
    """(photon numbers, ``ret[q, n]`` = gap of level q+1 to the ground state at n photons) (util.py:178-200)."""
    Erwa = np.asarray(Erwa)
    levels = Erwa.shape[0]
    photons = Erwa.shape[1] - levels
    q = np.arange(levels - 1)[:, None]
    n = np.arange(photons)[None, :]
    return np.arange(photons), Erwa[q + 1, n + q + 1] - Erwa[0, n]


def createSubspaceOperators(m00, m01, m10, m11):
    """Per-point 2x2 operators from four matrix-element sweeps (util.py:286-298), as one
    complex128 array [n_pts, 2, 2] (the reference builds a list of complex64 matrices)."""
    m = np.stack([np.asarray(m00), np.asarray(m01), np.asarray(m10), np.asarray(m11)], axis=-1)
    return np.ascontiguousarray(m.astype(np.complex128).reshape(-1, 2, 2))


def _as_vector(v):
    return np.asarray(v.full() if hasattr(v, "full") else v, dtype=np.complex128).reshape(-1)


def pauliCoefficients(E, V, basis_op, device=None):
    """Local-basis reduction to ``hx, hy, hz`` (util.py:300-391).

    ``E`` [levels, N] and ``V`` [levels, N] (eigenvector objects) as ``getSweep`` returns them with
    ``get_vectors=True``; ``basis_op`` is either one operator (anything with ``.full()`` / a dense
    array) or per-point 2x2 operators (list / array).  The 2x2 matrix elements of a fixed operator
    are formed on the device, and the per-point eigen-reduction runs as one kernel
    (``scqed_pauli_coefficients``) in float64 -- the reference loops in Python in complex64."""
    if not (type(E) in (list, np.ndarray) and (type(V) in (list, np.ndarray) or hasattr(V, "vectors"))):
        raise Exception("incompatible input types for E (%s), V (%s) and basis_op (%s)." % (type(E), type(V), type(basis_op)))
    E = np.asarray(E)
    E0 = np.asarray(E[0], dtype=np.float64)
    E1 = np.asarray(E[1], dtype=np.float64)
    n_pts = len(E0)
    if type(basis_op) in (list, np.ndarray) and np.asarray(basis_op).ndim == 3:
        Op = np.asarray(basis_op, dtype=np.complex128)
    elif hasattr(basis_op, "full") or (isinstance(basis_op, np.ndarray) and basis_op.ndim == 2):
        A = np.asarray(basis_op.full() if hasattr(basis_op, "full") else basis_op, dtype=np.complex128)
        if hasattr(V, "vectors") and hasattr(V, "shape"):           # lazy KetGrid from getSweep: no per-ket objects
            W = np.asarray(V[:2].vectors(), dtype=np.complex128).transpose(1, 0, 2)     # [N, 2, n]
        else:
            W = np.stack([np.stack([_as_vector(V[a][i]) for a in (0, 1)]) for i in range(n_pts)])
        Op = _engine.dense_matrix_elements(A, np.ascontiguousarray(W), device=device)   # Op[i, a, b] = <V_a| A |V_b>
    else:
        raise Exception("incompatible input types for E (%s), V (%s) and basis_op (%s)." % (type(E), type(V), type(basis_op)))
    if Op.shape != (n_pts, 2, 2):
        raise Exception("basis_op must give one 2x2 operator per sweep point")
    return _engine.pauli_coefficients(E0, E1, Op, device=device)


def getEigenValuesAndVectors(sweep):
    """(eigenvalues as float64, eigenvectors as objects) of a ``getSweep`` result with vectors (util.py:393-402)."""
    return (sweep[:, 0].astype(np.float64), sweep[:, 1])
