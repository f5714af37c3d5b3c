"""Per-node operator factors (the "factor table" of the sweep plan).

The reference builds six d x d operators per circuit node (Q, P, D, D^dagger,
S, S^dagger) and immediately Kronecker-expands each of them to the full
Hilbert space with ``qt.tensor`` (pyscqed/numerical_system.py:185-253).  Here
only the small per-node matrices are ever built; the engine keeps them
device-resident and forms  sum_t c_t(p) * kron_k O_{t,k}  on the fly.

This runs once per plan on the host (O(d^3) for one d x d ``eigh``); it is
plan construction, not the sweep loop.

Restated from (closed form where the reference goes through a numerical
eigendecomposition of an already-known operator):
  * charge basis      pyscqed/numerical_system.py:1111-1157
  * oscillator basis  pyscqed/numerical_system.py:1064-1109
  * flux basis        pyscqed/numerical_system.py:1159-1210
"""
from __future__ import annotations

import numpy as np

# pyscqed/physical_constants.py:5,11,29 (the reference hard-codes these CODATA-2014 values;
# they enter the oscillator displacement amplitudes, so they are reproduced verbatim).
E_CHARGE = 1.60217662e-19
HBAR = 1.054571800e-34
PHI0 = 2.067833831e-15

#: the reference thresholds every operator at 1e-12 (numerical_system.py:19,1103-1108,1151-1156)
QOBJ_ATOL = 1e-12

FACTOR_KINDS = ("charge", "flux", "disp", "disp_adj", "pdisp", "pdisp_adj")


def tidyup(a: np.ndarray, atol: float = QOBJ_ATOL) -> np.ndarray:
    """Zero real / imaginary parts below ``atol`` separately (QuTiP ``Qobj.tidyup`` semantics)."""
    a = np.array(a, dtype=np.complex128)
    re = a.real.copy()
    im = a.imag.copy()
    re[np.abs(re) < atol] = 0.0
    im[np.abs(im) < atol] = 0.0
    return re + 1j * im


def node_dimension(basis: str, trunc: int) -> int:
    """Hilbert-space dimension of one node (numerical_system.py:83-97; flux bases as 1159-1268)."""
    if basis == "oscillator":
        return int(trunc)
    if basis in ("charge", "flux", "discretized_flux"):
        return 2 * int(trunc) + 1
    raise Exception("Unrecognized basis representation '%s'." % repr(basis))


def _dft_conjugate(trunc: int) -> np.ndarray:
    """Dense Hermitian operator conjugate to the integer-ladder operator.

    The reference builds  P = sum_k (k/d) |phi_k><phi_k|  with
    <e_i|phi_k> = d^-1/2 exp(2 pi i k (t - i)/d), k = -t..t
    (numerical_system.py:1117-1131); in closed form
    P[a, b] = d^-2 sum_k k exp(2 pi i k (b - a)/d).
    """
    d = 2 * trunc + 1
    k = np.arange(-trunc, trunc + 1, dtype=np.float64)
    idx = np.arange(d)
    diff = (idx[None, :] - idx[:, None]).astype(np.float64)  # b - a
    phase = np.exp(2j * np.pi * k[None, None, :] * diff[:, :, None] / d)
    return (phase * k[None, None, :]).sum(axis=2) / (d * d)


def _shift_up(d: int) -> np.ndarray:
    """Ones on the first superdiagonal: exp(-2 pi i P) minus the wrap-around corner
    |2t><0| (numerical_system.py:1137-1145) is exactly this matrix."""
    return np.diag(np.ones(d - 1, dtype=np.complex128), 1)


def charge_basis(trunc: int):
    """(Q, P, D, Ddag, S, Sdag) in the charge basis, d = 2 trunc + 1.

    Q = diag(t, t-1, ..., -t) (numerical_system.py:1134); D raises the Cooper-pair number by
    one; S = exp(-2 pi i q) with integer q is the identity (numerical_system.py:1148).
    """
    d = 2 * trunc + 1
    Q = np.diag(np.arange(trunc, -trunc - 1, -1, dtype=np.float64)).astype(np.complex128)
    P = _dft_conjugate(trunc)
    D = _shift_up(d)
    S = np.eye(d, dtype=np.complex128)
    return tuple(tidyup(x) for x in (Q, P, D, D.conj().T, S, S.conj().T))


def flux_basis(trunc: int):
    """Dual of :func:`charge_basis` (numerical_system.py:1159-1210).

    P is the diagonal fluxon-number operator, Q the dense conjugate operator; the reference
    forms D = exp(-2 pi i P) - |2t><0| (P already diagonal and integer, so exp(...) is the
    identity) and S = exp(-2 pi i Q) - |2t><0| (the shift).
    """
    d = 2 * trunc + 1
    P = np.diag(np.arange(trunc, -trunc - 1, -1, dtype=np.float64)).astype(np.complex128)
    Q = _dft_conjugate(trunc)
    corner = np.zeros((d, d), dtype=np.complex128)
    corner[d - 1, 0] = 1.0
    D = np.eye(d, dtype=np.complex128) - corner
    S = _shift_up(d)
    return tuple(tidyup(x) for x in (Q, P, D, D.conj().T, S, S.conj().T))


def oscillator_ladder(trunc: int) -> np.ndarray:
    """Annihilation operator a, a[i, i+1] = sqrt(i+1) (QuTiP ``destroy``)."""
    return np.diag(np.sqrt(np.arange(1, trunc, dtype=np.float64)), 1).astype(np.complex128)


def oscillator_displacement_amplitudes(impedance_ohm: float, cooper_disp: float, fluxon_disp: float):
    """(beta, gamma): D = exp(i beta (a^dag + a)),  S = exp(i gamma i(a^dag - a)).

    From numerical_system.py:1076-1079 with the unit prefactors cancelled:
    beta  = cooper_disp * 2 pi sqrt(hbar) / Phi0 * sqrt(Z/2),
    gamma = fluxon_disp * pi sqrt(hbar) / e * sqrt(1/(2 Z)).
    """
    beta = cooper_disp * 2.0 * np.pi / PHI0 * np.sqrt(HBAR) * np.sqrt(impedance_ohm / 2.0)
    gamma = fluxon_disp * np.pi / E_CHARGE * np.sqrt(HBAR) * np.sqrt(1.0 / (2.0 * impedance_ohm))
    return beta, gamma


def _exp_i_hermitian(M: np.ndarray) -> np.ndarray:
    """exp(i M) for Hermitian M through its eigendecomposition (numerical_system.py:1082-1100)."""
    w, V = np.linalg.eigh(M)
    return (V * np.exp(1j * w)[None, :]) @ V.conj().T


def oscillator_basis(trunc: int, impedance_ohm: float, chg_osc: float, flx_osc: float,
                     cooper_disp: float = 1.0, fluxon_disp: float = 1.0):
    """(Q, P, D, Ddag, S, Sdag) in the Fock basis of the node's LC mode.

    ``impedance_ohm`` = Zosc<node> * Units.getPrefactor("Impe"); ``chg_osc``/``flx_osc`` are
    Units.getPrefactor("ChgOsc"/"FlxOsc") (numerical_system.py:1069-1077).
    """
    a = oscillator_ladder(trunc)
    ad = a.conj().T
    Q = 1j * np.sqrt(1.0 / (2.0 * impedance_ohm)) * (ad - a) * chg_osc
    P = np.sqrt(impedance_ohm / 2.0) * (ad + a) * flx_osc
    beta, gamma = oscillator_displacement_amplitudes(impedance_ohm, cooper_disp, fluxon_disp)
    D = _exp_i_hermitian(beta * (ad + a))
    S = _exp_i_hermitian(gamma * 1j * (ad - a))
    return tuple(tidyup(x) for x in (Q, P, D, D.conj().T, S, S.conj().T))


def oscillator_regen_set(trunc: int) -> dict:
    """Impedance-independent operator set of an oscillator node in the eigenbasis of X = a^dag + a.

    When the node impedance Z depends on a swept parameter the reference regenerates
    Q = i sqrt(1/2Z) (a^dag - a), P = sqrt(Z/2) (a^dag + a), D = exp(i beta(Z) X) at EVERY sweep point
    (numerical_system.py:378-393,435-437,1064-1109).  With X = W diag(lam) W^T (W real orthogonal) the
    similarity H' = W^T H W has, on this node, only CONSTANT matrices times scalar functions of Z:

        charge -> q(Z) W^T i(a^dag - a) W,   flux -> f(Z) diag(lam),   D -> diag(exp(i beta(Z) lam_m))

    so the Z dependence moves into the coefficient program (the diagonal D becomes one projector
    term per level) and eigenvectors are rotated back with W after the solve.
    Returns {'lam', 'W', 'charge', 'flux', 'charge2', 'flux2'} (unit matrices, d = trunc).
    """
    a = oscillator_ladder(trunc)
    ad = a.conj().T
    X = (ad + a).real
    lam, W = np.linalg.eigh(X)
    Qt = 1j * (ad - a)
    Qp = W.T @ Qt @ W
    Qp = 0.5 * (Qp + Qp.conj().T)
    return {"lam": lam, "W": W, "charge": Qp, "flux": np.diag(lam).astype(np.complex128),
            "charge2": Qp @ Qp, "flux2": np.diag(lam * lam).astype(np.complex128)}


def classify(M: np.ndarray) -> str:
    """Structural type of a factor: 'identity', 'diag', 'band1' (|i-j| <= 1) or 'dense'.

    The engine uses this to pick structure-aware code paths (e.g. a Hamiltonian whose every
    term is band1 on a single node is Hermitian tridiagonal and skips the Householder stage).
    """
    d = M.shape[0]
    if np.array_equal(M, np.eye(d)):
        return "identity"
    off = M - np.diag(np.diag(M))
    if not off.any():
        return "diag"
    i, j = np.nonzero(off)
    if np.max(np.abs(i - j)) <= 1:
        return "band1"
    return "dense"
