"""Minimal QuTiP-compatibility module -- TEST INFRASTRUCTURE ONLY (oracle side).

QuTiP 5.0.3.post1 (the reference's pinned dependency, poetry.lock:2636) is not
installed in this image and cannot be installed offline.  This module provides
exactly the QuTiP surface that the reference's hot path touches
(call sites pyscqed/numerical_system.py:100,109-123,
195-251,594,666,693,755-759,1076-1267 and pyscqed/util.py:96-122,147-173) so
that the UNMODIFIED reference package can be imported from /root/reference to
mint golden vectors (tests/golden/make_golden.py).

It restates QuTiP's published semantics on top of scipy.sparse CSR:
  * Qobj * Qobj is a matrix product, Qobj * scalar scales;
  * Qobj + scalar adds scalar * identity (Qobj.__add__ in qutip/core/qobj.py);
  * tidyup(atol) zeroes real and imaginary parts separately below atol;
  * tensor() is the Kronecker product, first argument = slowest index;
  * eigenstates()/eigenenergies() are dense Hermitian solves, ascending.

Nothing in the product path (pycqed_b200/) may import this module.
"""
import numbers

import numpy as np
import scipy.sparse as sp

__version__ = "5.0.3.post1-shim"


def _is_scalar(x):
    return isinstance(x, (numbers.Number, np.number)) or (
        isinstance(x, np.ndarray) and x.ndim == 0
    )


class _Data:
    """Stand-in for qutip.core.data.{CSR,Dense}: to_array() / as_scipy()."""

    def __init__(self, mat):
        self._mat = mat

    def to_array(self):
        return np.asarray(self._mat.toarray(), dtype=np.complex128)

    def as_scipy(self):
        return self._mat


class Qobj:
    __array_ufunc__ = None  # make numpy scalars / arrays defer to our operators
    __array_priority__ = 100

    def __init__(self, arg=None, dims=None, copy=True):
        if isinstance(arg, Qobj):
            mat = arg._mat.copy()
            if dims is None:
                dims = [list(arg.dims[0]), list(arg.dims[1])]
        elif sp.issparse(arg):
            mat = sp.csr_matrix(arg, dtype=np.complex128)
        else:
            a = np.asarray(arg, dtype=np.complex128)
            if a.ndim == 0:
                a = a.reshape(1, 1)
            elif a.ndim == 1:
                a = a.reshape(-1, 1)  # 1-D input is a ket
            mat = sp.csr_matrix(a)
        self._mat = mat
        if dims is None:
            dims = [[mat.shape[0]], [mat.shape[1]]]
        self.dims = [list(dims[0]), list(dims[1])]

    # -- structure ---------------------------------------------------------
    @property
    def data(self):
        return _Data(self._mat)

    @property
    def shape(self):
        return self._mat.shape

    @property
    def isket(self):
        return self._mat.shape[1] == 1 and self._mat.shape[0] > 1

    @property
    def isbra(self):
        return self._mat.shape[0] == 1 and self._mat.shape[1] > 1

    @property
    def isoper(self):
        return self._mat.shape[0] == self._mat.shape[1]

    @property
    def isherm(self):
        if self._mat.shape[0] != self._mat.shape[1]:
            return False
        d = self._mat - self._mat.conj().T
        if d.nnz == 0:
            return True
        return bool(np.max(np.abs(d.data)) <= 1e-12)

    def to(self, _fmt):
        return self

    def copy(self):
        return Qobj(self._mat.copy(), dims=self.dims)

    def full(self):
        return self.data.to_array()

    def dag(self):
        return Qobj(self._mat.conj().T.tocsr(), dims=[self.dims[1], self.dims[0]])

    def tidyup(self, atol=1e-14):
        m = self._mat.tocsr()
        re = m.data.real.copy()
        im = m.data.imag.copy()
        re[np.abs(re) < atol] = 0.0
        im[np.abs(im) < atol] = 0.0
        m.data = re + 1j * im
        m.eliminate_zeros()
        self._mat = m
        return self

    def norm(self):
        if self.isket or self.isbra:
            return float(np.sqrt(np.sum(np.abs(self._mat.data) ** 2)))
        # trace norm for operators
        return float(np.sum(np.linalg.svd(self.full(), compute_uv=False)))

    def tr(self):
        return self._mat.diagonal().sum()

    def matrix_element(self, bra, ket):
        left = bra._mat
        if bra.isket:
            left = left.conj().T
        out = (left @ (self._mat @ ket._mat)).toarray()
        return complex(out[0, 0])

    def eigenstates(self):
        w, v = np.linalg.eigh(self.full())
        kets = np.empty(len(w), dtype=object)
        for i in range(len(w)):
            kets[i] = Qobj(v[:, i], dims=[self.dims[0], [1] * len(self.dims[0])])
        return w, kets

    def eigenenergies(self):
        return np.linalg.eigvalsh(self.full())

    # -- arithmetic --------------------------------------------------------
    def _eye_like(self):
        return sp.identity(self._mat.shape[0], dtype=np.complex128, format="csr")

    def __add__(self, other):
        if isinstance(other, Qobj):
            return Qobj(self._mat + other._mat, dims=self.dims)
        if _is_scalar(other):
            if other == 0:
                return self.copy()
            return Qobj(self._mat + complex(other) * self._eye_like(), dims=self.dims)
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, Qobj):
            return Qobj(self._mat - other._mat, dims=self.dims)
        if _is_scalar(other):
            return self + (-other)
        return NotImplemented

    def __rsub__(self, other):
        return (-self) + other

    def __neg__(self):
        return Qobj(-self._mat, dims=self.dims)

    def __mul__(self, other):
        if isinstance(other, Qobj):
            return Qobj((self._mat @ other._mat).tocsr(), dims=[self.dims[0], other.dims[1]])
        if _is_scalar(other):
            return Qobj(self._mat * complex(other), dims=self.dims)
        return NotImplemented

    def __matmul__(self, other):
        return self.__mul__(other)

    def __rmul__(self, other):
        if _is_scalar(other):
            return Qobj(self._mat * complex(other), dims=self.dims)
        return NotImplemented

    def __truediv__(self, other):
        if _is_scalar(other):
            return Qobj(self._mat / complex(other), dims=self.dims)
        return NotImplemented

    def __eq__(self, other):
        if not isinstance(other, Qobj):
            return False
        if self.shape != other.shape:
            return False
        d = self._mat - other._mat
        return d.nnz == 0 or bool(np.max(np.abs(d.data)) <= 1e-12)

    __hash__ = None

    def __repr__(self):
        return "Qobj(shim) dims=%r shape=%r\n%r" % (self.dims, self.shape, self.full())


class _QobjModule:
    """So that `qt.qobj.Qobj` resolves (pyscqed/util.py:96,147)."""

    Qobj = Qobj


qobj = _QobjModule()


def qeye(N):
    return Qobj(sp.identity(int(N), dtype=np.complex128, format="csr"))


identity = qeye


def num(N, offset=0):
    return Qobj(sp.diags(np.arange(offset, offset + N, dtype=np.complex128), format="csr"))


def basis(N, n=0):
    v = np.zeros(int(N), dtype=np.complex128)
    v[int(n)] = 1.0
    return Qobj(v)


def destroy(N):
    return Qobj(sp.diags(np.sqrt(np.arange(1, N, dtype=np.complex128)), 1, format="csr"))


def create(N):
    return Qobj(sp.diags(np.sqrt(np.arange(1, N, dtype=np.complex128)), -1, format="csr"))


def tensor(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple, np.ndarray)):
        args = list(args[0])
    out = args[0]._mat
    d0 = list(args[0].dims[0])
    d1 = list(args[0].dims[1])
    for a in args[1:]:
        out = sp.kron(out, a._mat, format="csr")
        d0 += list(a.dims[0])
        d1 += list(a.dims[1])
    return Qobj(out, dims=[d0, d1])


def commutator(A, B, kind="normal"):
    if kind == "normal":
        return A * B - B * A
    return A * B + B * A


def sigmax():
    return Qobj(np.array([[0, 1], [1, 0]], dtype=np.complex128))


def sigmay():
    return Qobj(np.array([[0, -1j], [1j, 0]], dtype=np.complex128))


def sigmaz():
    return Qobj(np.array([[1, 0], [0, -1]], dtype=np.complex128))
