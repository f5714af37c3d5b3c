"""ctypes binding of ``libscqed.so`` (include/scqed.h) -- the only way numbers are produced.

There is NO CPU fallback: if the CUDA library is missing or no CUDA device is visible, every
compute entry point raises.  PyTorch is used for device buffers and streams only.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from .densetab import build_dense_table
from .entrytab import build_entry_table

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscqed.so")

SOLVER_AUTO, SOLVER_SMALL, SOLVER_DENSE, SOLVER_DENSE_UNBLOCKED, SOLVER_MATFREE, SOLVER_SMALL_FUSED = 0, 1, 2, 3, 4, 5


class EngineError(RuntimeError):
    pass


class _PlanDesc(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32), ("dims", C.POINTER(C.c_int32)),
        ("n_factors", C.c_int32), ("factor_node", C.POINTER(C.c_int32)),
        ("factor_offset", C.POINTER(C.c_int64)), ("factor_data", C.POINTER(C.c_double)),
        ("n_terms", C.c_int32), ("term_factors", C.POINTER(C.c_int32)),
        ("n_swept", C.c_int32), ("n_instr", C.c_int32), ("instr", C.POINTER(C.c_int32)),
        ("n_consts", C.c_int32), ("consts", C.POINTER(C.c_double)),
        ("coef_re", C.POINTER(C.c_int32)), ("coef_im", C.POINTER(C.c_int32)),
        ("n_scalar", C.c_int32), ("scalar_regs", C.POINTER(C.c_int32)),
        ("n_aux", C.c_int32), ("aux_term_factors", C.POINTER(C.c_int32)),
        ("aux_re", C.POINTER(C.c_int32)), ("aux_im", C.POINTER(C.c_int32)),
        ("n_pairs", C.c_int32), ("pairs", C.POINTER(C.c_int32)),
        ("n_dense", C.c_int32), ("n_dense_diag", C.c_int32), ("dense_tridiag", C.c_int32), ("dense_tab", C.POINTER(C.c_double)),
        ("dense_re", C.POINTER(C.c_int32)), ("dense_im", C.POINTER(C.c_int32)), ("dense_gmax", C.POINTER(C.c_double)),
        ("n_entries", C.c_int32), ("ent_dense", C.c_int32), ("ent_key", C.POINTER(C.c_int64)),
        ("ent_ptr", C.POINTER(C.c_int32)), ("ent_term", C.POINTER(C.c_int32)), ("ent_val", C.POINTER(C.c_double)),
        ("n_xform", C.c_int32), ("xform_node", C.POINTER(C.c_int32)), ("xform_data", C.POINTER(C.c_double)),
    ]


class _SweepOpts(C.Structure):
    _fields_ = [
        ("k", C.c_int32), ("want_vectors", C.c_int32), ("solver", C.c_int32), ("tidyup", C.c_int32),
        ("resonator", C.c_int32), ("res_nmax", C.c_int32), ("res_op_factor", C.c_int32),
        ("res_gc", C.c_int32), ("res_frl", C.c_int32), ("res_qb", C.c_int32), ("raw_basis", C.c_int32),
        ("real_vectors", C.c_int32), ("evecs_on_device", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]


#: every symbol include/scqed.h declares (tests check that the library exports all of them)
EXPORTS = (
    "scqed_plan_create", "scqed_plan_destroy", "scqed_plan_hilbert_dim", "scqed_eval_coefficients",
    "scqed_assemble_dense", "scqed_eigh_batched", "scqed_sweep_run", "scqed_sweep_run_host", "scqed_matrix_elements", "scqed_pauli_coefficients", "scqed_back_transform",
    "scqed_dense_matrix_elements",
    "scqed_launch_count", "scqed_timing_enable", "scqed_timing_read", "scqed_counter_read", "scqed_probe_fp64_tensor", "scqed_probe_fp64_fma", "scqed_last_error",
    "scqed_version",
)

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """Load ``libscqed.so`` (built in-tree by ``__graft_entry__.build()``); raises if absent."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise EngineError(
                "CUDA extension %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
        lib.scqed_plan_create.argtypes = [C.POINTER(_PlanDesc), C.c_int, C.POINTER(vp)]
        lib.scqed_plan_create.restype = C.c_int
        lib.scqed_plan_destroy.argtypes = [vp]
        lib.scqed_plan_destroy.restype = None
        lib.scqed_plan_hilbert_dim.argtypes = [vp]
        lib.scqed_plan_hilbert_dim.restype = C.c_int
        lib.scqed_eval_coefficients.argtypes = [vp, dp, i64, dp, dp, vp]
        lib.scqed_eval_coefficients.restype = C.c_int
        lib.scqed_assemble_dense.argtypes = [vp, dp, i64, i32, dp, vp]
        lib.scqed_assemble_dense.restype = C.c_int
        lib.scqed_eigh_batched.argtypes = [C.c_int, dp, i32, i64, i32, dp, dp, dp, i32, vp]
        lib.scqed_eigh_batched.restype = C.c_int
        lib.scqed_sweep_run.argtypes = [vp, C.POINTER(_SweepOpts), dp, i64, dp, dp, dp, dp, dp, vp]
        lib.scqed_sweep_run.restype = C.c_int
        lib.scqed_sweep_run_host.argtypes = [vp, C.POINTER(_SweepOpts), dp, i64, dp, dp, dp, dp, dp]
        lib.scqed_sweep_run_host.restype = C.c_int
        lib.scqed_matrix_elements.argtypes = [vp, dp, i64, i32, dp, dp, vp]
        lib.scqed_matrix_elements.restype = C.c_int
        lib.scqed_back_transform.argtypes = [vp, dp, i64, vp]
        lib.scqed_back_transform.restype = C.c_int
        lib.scqed_pauli_coefficients.argtypes = [i32, dp, dp, i64, dp, vp]
        lib.scqed_pauli_coefficients.restype = C.c_int
        lib.scqed_dense_matrix_elements.argtypes = [i32, dp, dp, i32, i32, i64, dp, vp]
        lib.scqed_dense_matrix_elements.restype = C.c_int
        lib.scqed_launch_count.argtypes = []
        lib.scqed_launch_count.restype = C.c_int64
        lib.scqed_timing_enable.argtypes = [C.c_int]
        lib.scqed_timing_enable.restype = None
        lib.scqed_timing_read.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        lib.scqed_timing_read.restype = C.c_int
        lib.scqed_counter_read.argtypes = [C.c_int, C.POINTER(C.c_int64)]
        lib.scqed_counter_read.restype = C.c_int
        lib.scqed_probe_fp64_tensor.argtypes = [C.c_int, C.POINTER(C.c_double)]
        lib.scqed_probe_fp64_tensor.restype = C.c_int
        lib.scqed_probe_fp64_fma.argtypes = [C.c_int, C.POINTER(C.c_double)]
        lib.scqed_probe_fp64_fma.restype = C.c_int
        lib.scqed_last_error.argtypes = []
        lib.scqed_last_error.restype = C.c_char_p
        lib.scqed_version.argtypes = []
        lib.scqed_version.restype = C.c_char_p
        _lib = lib
        return lib


def _check(rc):
    if rc != 0:
        msg = load_library().scqed_last_error().decode("utf-8", "replace")
        raise EngineError("libscqed error %d: %s" % (rc, msg))


def _require_cuda(device=None):
    import torch
    if not torch.cuda.is_available():
        raise EngineError("no CUDA device visible; the sweep engine has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device() if device is None else int(device))


def launch_count() -> int:
    return int(load_library().scqed_launch_count())


def timing_enable(on=True):
    load_library().scqed_timing_enable(1 if on else 0)


def timing_read(tag):
    """(total ms, launches) of the timed kernel spans with this tag since the last read."""
    ms, cnt = C.c_double(), C.c_int64()
    _check(load_library().scqed_timing_read(int(tag), C.byref(ms), C.byref(cnt)))
    return ms.value, cnt.value


def counter_read(which):
    """Work counter ``which`` since the last read (1: vector mat-vecs of the matrix-free filters)."""
    v = C.c_int64()
    _check(load_library().scqed_counter_read(int(which), C.byref(v)))
    return int(v.value)


def probe_fp64(device=0):
    """(DMMA TFLOP/s, DFMA TFLOP/s) measured with register-resident loops."""
    lib = load_library()
    _require_cuda(device)
    a, b = C.c_double(), C.c_double()
    _check(lib.scqed_probe_fp64_tensor(int(device), C.byref(a)))
    _check(lib.scqed_probe_fp64_fma(int(device), C.byref(b)))
    return a.value, b.value


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))


def _ptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def _stream_ptr(device=None):
    """cudaStream_t of torch's current stream ON ``device`` (not on whatever device happens to be current)."""
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def pinned_empty(shape, dtype):
    """A numpy array backed by page-locked host memory (torch's caching host allocator: the block is reused by
    later sweeps once this array is garbage collected).  D2H copies into it are true asynchronous DMA."""
    import torch
    tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.complex128): torch.complex128,
           np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
    return torch.empty(tuple(int(x) for x in shape), dtype=tdt, pin_memory=True).numpy()


def _check_out(name, arr, shape, dtypes):
    """Caller-supplied result buffers are handed to the library by raw pointer: refuse anything that is not exactly
    the array the library is going to write."""
    if not isinstance(arr, np.ndarray):
        raise EngineError("out[%r] must be a numpy array" % name)
    if arr.dtype not in [np.dtype(d) for d in dtypes]:
        raise EngineError("out[%r] has dtype %s, expected %s" % (name, arr.dtype, " or ".join(str(np.dtype(d)) for d in dtypes)))
    if tuple(arr.shape) != tuple(shape):
        raise EngineError("out[%r] has shape %r, expected %r" % (name, tuple(arr.shape), tuple(shape)))
    if not arr.flags.c_contiguous or not arr.flags.writeable:
        raise EngineError("out[%r] must be C-contiguous and writeable" % name)
    return arr


class SweepResult:
    """Results of one sweep on the host (numpy).  ``E`` [n_pts, k]; ``V`` [n_pts, k, n] or None;
    ``scalars`` {name: [n_pts]}; ``Erwa`` [n_pts, nmax+1+k, k] or None; ``info`` [n_pts].
    ``V_dev``: when the sweep was asked to keep the eigenvectors on the device, the list of
    ``(lo, hi, tensor[hi-lo, k, n])`` shards (one per GPU) and ``V`` is fetched lazily by :meth:`vectors`."""

    def __init__(self, E, V, scalars, Erwa, info, matel=None, V_dev=None):
        self.E, self.V, self.scalars, self.Erwa, self.info = E, V, scalars, Erwa, info
        self.matel = matel          # complex [n_pts, n_pairs] (Current / Voltage matrix elements) or None
        self.V_dev = V_dev

    def vectors(self):
        """Host eigenvectors [n_pts, k, n]; copies the device shards on first use."""
        if self.V is None and self.V_dev:
            n_pts = max(hi for _, hi, _ in self.V_dev)
            t0 = self.V_dev[0][2]
            V = pinned_empty((n_pts,) + tuple(t0.shape[1:]), np.complex128)
            import torch
            Vt = torch.from_numpy(V)
            for lo, hi, t in self.V_dev:
                Vt[lo:hi].copy_(t, non_blocking=True)
            for _, _, t in self.V_dev:
                torch.cuda.synchronize(t.device)
            self.V = V
        return self.V


class Engine:
    """A :class:`~pycqed_b200.plan.SweepPlan` uploaded to one GPU."""

    def __init__(self, plan, device=None):
        self.lib = load_library()
        self.device = _require_cuda(device)
        self.plan = plan
        self.n = plan.hilbert_dim
        self._handle = C.c_void_p()
        prog = plan.program
        # small Hilbert spaces: dense real term table for the fused solver (extends the program)
        dense = None if os.environ.get("SCQED_NO_DENSE_TABLE") else build_dense_table(plan)
        if dense is not None:
            prog = dense["program"]
        dims = _i32(plan.dims)
        fnode = _i32([f[0] for f in plan.factors])
        mats = [np.ascontiguousarray(plan.factor_matrix(i), dtype=np.complex128) for i in range(len(plan.factors))]
        offs, off = [], 0
        for m in mats:
            offs.append(off)
            off += m.size
        foff = np.ascontiguousarray(np.asarray(offs, dtype=np.int64))
        fdata = np.concatenate([m.reshape(-1) for m in mats]) if mats else np.zeros(1, dtype=np.complex128)
        fdata = np.ascontiguousarray(fdata).view(np.float64)
        terms = _i32(plan.terms)
        instr = _i32(prog.instr)
        consts = np.ascontiguousarray(np.asarray(prog.consts, dtype=np.float64))
        if consts.size == 0:
            consts = np.zeros(1, dtype=np.float64)
        cre, cim, sreg = _i32(prog.coef_re), _i32(prog.coef_im), _i32(prog.scalar_regs)
        if sreg.size == 0:
            sreg = np.zeros(1, dtype=np.int32)
        aux_terms = _i32(plan.aux_terms) if plan.aux_terms else np.zeros(1, dtype=np.int32)
        aux_re = _i32(prog.aux_re) if len(prog.aux_re) else np.zeros(1, dtype=np.int32)
        aux_im = _i32(prog.aux_im) if len(prog.aux_im) else np.zeros(1, dtype=np.int32)
        pair_list = plan.aux_pairs()
        pairs = _i32(pair_list) if pair_list else np.zeros(4, dtype=np.int32)
        self.n_pairs = len(pair_list)
        if len(plan.aux_terms) != len(prog.aux_re):
            raise EngineError("plan is inconsistent: %d aux terms but %d aux coefficients" % (len(plan.aux_terms), len(prog.aux_re)))
        if len(plan.terms) != prog.n_coef:
            raise EngineError("plan is inconsistent: %d terms but %d coefficients" % (len(plan.terms), prog.n_coef))
        desc = _PlanDesc(
            n_nodes=len(plan.dims), dims=_ptr(dims, C.c_int32),
            n_factors=len(mats), factor_node=_ptr(fnode, C.c_int32),
            factor_offset=_ptr(foff, C.c_int64), factor_data=_ptr(fdata, C.c_double),
            n_terms=len(plan.terms), term_factors=_ptr(terms, C.c_int32),
            n_swept=prog.n_swept, n_instr=prog.n_instr, instr=_ptr(instr, C.c_int32),
            n_consts=len(prog.consts), consts=_ptr(consts, C.c_double),
            coef_re=_ptr(cre, C.c_int32), coef_im=_ptr(cim, C.c_int32),
            n_scalar=prog.n_scalar, scalar_regs=_ptr(sreg, C.c_int32),
            n_aux=len(plan.aux_terms), aux_term_factors=_ptr(aux_terms, C.c_int32),
            aux_re=_ptr(aux_re, C.c_int32), aux_im=_ptr(aux_im, C.c_int32),
            n_pairs=len(pair_list), pairs=_ptr(pairs, C.c_int32),
        )
        # dense assembly (K2, also feeding the HBM dense solver): entry list of the pattern of H.
        # Only where a dense H is a sane object (<= 2 GB per matrix).
        ent = None
        if self.n <= 11000 and not os.environ.get("SCQED_NO_ENTRY_LIST"):
            ent = build_entry_table(plan)
        if ent is not None:
            ekey = np.ascontiguousarray(ent["key"], dtype=np.int64)
            eptr, eterm = _i32(ent["ptr"]), _i32(ent["term"])
            evalv = np.ascontiguousarray(ent["val"], dtype=np.complex128).view(np.float64)
            desc.n_entries, desc.ent_dense = int(ekey.size), int(ent["dense"])
            desc.ent_key = _ptr(ekey, C.c_int64)
            desc.ent_ptr, desc.ent_term, desc.ent_val = _ptr(eptr, C.c_int32), _ptr(eterm, C.c_int32), _ptr(evalv, C.c_double)
        self.tridiagonal = bool(dense is not None and dense["tridiagonal"] and not os.environ.get("SCQED_NO_TRIDIAG_SHORTCUT"))
        xf = plan.node_transforms()
        self.rotated = bool(xf)
        if xf:
            xnode = _i32(sorted(xf))
            xdata = np.ascontiguousarray(np.concatenate([np.asarray(xf[i], dtype=np.float64).reshape(-1) for i in sorted(xf)]))
            desc.n_xform = len(xf)
            desc.xform_node, desc.xform_data = _ptr(xnode, C.c_int32), _ptr(xdata, C.c_double)
        if dense is not None:
            dtab = np.ascontiguousarray(dense["table"], dtype=np.float64)
            dre, dim_, dgm = _i32(dense["re_regs"]), _i32(dense["im_regs"]), np.ascontiguousarray(dense["gmax"], dtype=np.float64)
            desc.n_dense = int(dense["n_full"] + dense["n_diag"])
            desc.n_dense_diag = int(dense["n_diag"])
            desc.dense_tridiag = int(dense["tridiagonal"] and not os.environ.get("SCQED_NO_TRIDIAG_SHORTCUT"))
            desc.dense_tab = _ptr(dtab, C.c_double)
            desc.dense_re, desc.dense_im = _ptr(dre, C.c_int32), _ptr(dim_, C.c_int32)
            desc.dense_gmax = _ptr(dgm, C.c_double)
        _check(self.lib.scqed_plan_create(C.byref(desc), self.device.index, C.byref(self._handle)))

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self.lib.scqed_plan_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _st(self):
        return _stream_ptr(self.device)

    # ------------------------------------------------------------------------------------------
    def _opts(self, k, want_vectors, solver, tidyup, resonator):
        o = _SweepOpts()
        o.k, o.want_vectors, o.solver, o.tidyup = int(k), int(bool(want_vectors)), int(solver), int(bool(tidyup))
        if resonator:
            spec = self.plan.post.get("resonator")
            if spec is None:
                raise EngineError("this plan was lowered without the Resonator evaluable")
            o.resonator = 1
            o.res_nmax = int(spec["nmax"])
            o.res_op_factor = int(spec["op_factor"])
            o.res_gc, o.res_frl, o.res_qb = int(spec["gC"]), int(spec["frl"]), int(spec["qb"])
        return o

    def _params_dev(self, params):
        import torch
        if isinstance(params, torch.Tensor):
            p = params.to(device=self.device, dtype=torch.float64).contiguous()
        else:
            p = torch.as_tensor(np.ascontiguousarray(params, dtype=np.float64), device=self.device)
        if p.ndim != 2 or p.shape[1] != self.plan.program.n_swept:
            raise EngineError("params must be [n_pts, %d]" % self.plan.program.n_swept)
        return p

    def coefficients(self, params):
        """K1 on the device: (coef complex [n_pts, n_terms], scalars [n_pts, n_scalar]) as torch tensors."""
        import torch
        p = self._params_dev(params)
        n_pts = p.shape[0]
        coef = torch.empty((n_pts, len(self.plan.terms)), dtype=torch.complex128, device=self.device)
        scal = torch.empty((n_pts, max(self.plan.program.n_scalar, 1)), dtype=torch.float64, device=self.device)
        _check(self.lib.scqed_eval_coefficients(self._handle, p.data_ptr(), n_pts, coef.data_ptr(), scal.data_ptr(), self._st()))
        return coef, scal[:, :self.plan.program.n_scalar]

    def assemble_dense(self, params, tidyup=True):
        """K2: dense H [n_pts, n, n] complex128 torch tensor on the device."""
        import torch
        p = self._params_dev(params)
        n_pts = p.shape[0]
        H = torch.empty((n_pts, self.n, self.n), dtype=torch.complex128, device=self.device)
        _check(self.lib.scqed_assemble_dense(self._handle, p.data_ptr(), n_pts, int(bool(tidyup)), H.data_ptr(), self._st()))
        return H

    def matrix_elements(self, params, V):
        """<i|Op|j> of the plan's auxiliary operators (Current / Voltage evaluables) for every
        point; ``V`` is the device tensor [n_pts, k, n] returned by :meth:`sweep_device` (for a plan with
        rotated nodes: by ``sweep_device(..., raw_basis=True)``, i.e. in the plan's basis).  Returns a
        complex128 device tensor [n_pts, n_pairs] ordered like ``plan.aux_pairs()``."""
        import torch
        if self.n_pairs < 1:
            raise EngineError("this plan was lowered without matrix-element evaluations")
        p = self._params_dev(params)
        n_pts = p.shape[0]
        if isinstance(V, np.ndarray):
            V = torch.as_tensor(np.ascontiguousarray(V, dtype=np.complex128), device=self.device)
        if not (isinstance(V, torch.Tensor) and V.is_cuda and V.dtype == torch.complex128 and V.ndim == 3
                and V.shape[0] == n_pts and V.shape[2] == self.n):
            raise EngineError("V must be a complex128 array / CUDA tensor [n_pts, k, n] of eigenvectors")
        V = V.contiguous()
        out = torch.empty((n_pts, self.n_pairs), dtype=torch.complex128, device=self.device)
        _check(self.lib.scqed_matrix_elements(self._handle, p.data_ptr(), n_pts, int(V.shape[1]), V.data_ptr(), out.data_ptr(), self._st()))
        return out

    def back_transform(self, V):
        """Rotate eigenvectors [n_pts, k, n] (CUDA complex128, in place) from the plan's basis to the
        reference's Fock basis (only plans with swept-impedance oscillator nodes have one)."""
        if self.rotated:
            _check(self.lib.scqed_back_transform(self._handle, V.data_ptr(), int(V.shape[0] * V.shape[1]), self._st()))
        return V

    def sweep_device(self, params, k, want_vectors=False, resonator=False, solver=SOLVER_AUTO, tidyup=True,
                     raw_basis=False):
        """Device-resident sweep; returns torch tensors (E, V|None, scalars, Erwa|None, info)."""
        import torch
        p = self._params_dev(params)
        n_pts = p.shape[0]
        o = self._opts(k, want_vectors or resonator, solver, tidyup, resonator)
        o.raw_basis = int(bool(raw_basis))
        E = torch.empty((n_pts, k), dtype=torch.float64, device=self.device)
        V = torch.empty((n_pts, k, self.n), dtype=torch.complex128, device=self.device) if want_vectors else None
        ns = self.plan.program.n_scalar
        scal = torch.empty((n_pts, max(ns, 1)), dtype=torch.float64, device=self.device)
        post = None
        if resonator:
            post = torch.empty((n_pts, o.res_nmax + 1 + k, k), dtype=torch.float64, device=self.device)
        info = torch.empty((n_pts,), dtype=torch.int32, device=self.device)
        _check(self.lib.scqed_sweep_run(
            self._handle, C.byref(o), p.data_ptr(), n_pts, E.data_ptr(), V.data_ptr() if V is not None else None,
            scal.data_ptr(), post.data_ptr() if post is not None else None, info.data_ptr(), self._st()))
        return E, V, scal[:, :ns], post, info

    def sweep(self, params, k, want_vectors=False, resonator=False, solver=SOLVER_AUTO, tidyup=True,
              out=None, matrix_elements=False, real_vectors=False, vectors_on_device=False, row_offset=0) -> SweepResult:
        """Host-buffer sweep (the call the drop-in ``paramSweep`` makes): params and results are numpy arrays, the
        library does the H2D / D2H copies (``scqed_sweep_run_host``).

        Result arrays the caller does not supply in ``out`` are allocated PAGE-LOCKED (:func:`pinned_empty`), so the
        copies out of the GPU are asynchronous DMA overlapped with the next chunk's solve; pageable arrays supplied
        in ``out`` work too (the library stages them through its own pinned ring).  ``out`` = {'E','V','Erwa',
        'scalars','info'}: arrays of the FULL result; this call writes rows [row_offset, row_offset + n_pts) (how
        several GPUs fill one result).

        ``matrix_elements``: the eigenvectors stay on the device for ``scqed_matrix_elements`` and everything is
        copied back afterwards.  ``real_vectors``: True / "auto" asks the library to ship only the real parts of the
        eigenvectors (``V`` comes back float64 [n_pts, k, n]: a third less PCIe traffic for real symmetric H); if
        some eigenvector is complex, "auto" repeats the call with a complex buffer and True raises.
        ``vectors_on_device``: eigenvectors are NOT copied to the host; ``SweepResult.V_dev`` holds the device
        tensor and ``SweepResult.vectors()`` fetches them on demand."""
        import torch
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.ndim != 2 or params.shape[1] != self.plan.program.n_swept:
            raise EngineError("params must be [n_pts, %d]" % self.plan.program.n_swept)
        n_pts = params.shape[0]
        names = list(self.plan.program.scalar_names)
        with torch.cuda.device(self.device):
            if matrix_elements:
                p = self._params_dev(params)
                E, V, scal, post, info = self.sweep_device(p, k, want_vectors=True, resonator=resonator, solver=solver,
                                                           tidyup=tidyup, raw_basis=True)
                M = self.matrix_elements(p, V)            # operators and vectors both in the plan's basis
                self.back_transform(V)
                scal = scal.cpu().numpy()
                keep = want_vectors and vectors_on_device
                return SweepResult(E.cpu().numpy(), V.cpu().numpy() if (want_vectors and not keep) else None,
                                   {nm: scal[:, i].copy() for i, nm in enumerate(names)},
                                   post.cpu().numpy() if post is not None else None, info.cpu().numpy(), M.cpu().numpy(),
                                   V_dev=[(row_offset, row_offset + n_pts, V)] if keep else None)
            o = self._opts(k, want_vectors or resonator, solver, tidyup, resonator)
            ns = self.plan.program.n_scalar
            out = out or {}
            r0 = int(row_offset)

            def rows(name, tail, dtypes):
                """The [n_pts, ...] block of out[name] this call writes (validated), or a fresh pinned array."""
                full = out.get(name)
                if full is None:
                    return pinned_empty((n_pts,) + tail, dtypes[0])
                if full.ndim < 1 or full.shape[0] < r0 + n_pts:
                    raise EngineError("out[%r] has %d rows, this call writes rows %d..%d" % (name, full.shape[0] if full.ndim else 0, r0, r0 + n_pts))
                _check_out(name, full, (full.shape[0],) + tail, dtypes)
                return full[r0:r0 + n_pts]

            E = rows("E", (k,), [np.float64])
            V = V_dev = None
            to_dev = bool(vectors_on_device) and want_vectors
            real_req = bool(real_vectors) and want_vectors and not self.rotated and not to_dev
            if want_vectors and not to_dev:
                if out.get("V") is not None:
                    V = rows("V", (k, self.n), [np.complex128, np.float64])
                    real_req = real_req and V.dtype == np.float64
                    if V.dtype == np.float64 and not real_req:
                        raise EngineError("out['V'] is float64 but real_vectors was not requested (or the plan has rotated nodes)")
                else:
                    V = pinned_empty((n_pts, k, self.n), np.float64 if real_req else np.complex128)
            if to_dev:
                V_dev = torch.empty((n_pts, k, self.n), dtype=torch.complex128, device=self.device)
                o.evecs_on_device = 1
            o.real_vectors = int(real_req)
            scal = rows("scalars", (max(ns, 1),), [np.float64])
            post = rows("Erwa", (o.res_nmax + 1 + k, k), [np.float64]) if resonator else None
            info = rows("info", (), [np.int32])
            vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
            vptr = C.c_void_p(V_dev.data_ptr()) if to_dev else vp(V)
            rc = self.lib.scqed_sweep_run_host(self._handle, C.byref(o), vp(params), n_pts, vp(E), vptr, vp(scal),
                                               vp(post), vp(info))
            if rc == 1 and real_req:                 # SCQED_ECOMPLEX: the eigenvectors are not real
                if real_vectors != "auto" or out.get("V") is not None:
                    raise EngineError("real_vectors=True but the eigenvectors are complex; use real_vectors='auto' or False")
                o.real_vectors = 0
                V = pinned_empty((n_pts, k, self.n), np.complex128)
                rc = self.lib.scqed_sweep_run_host(self._handle, C.byref(o), vp(params), n_pts, vp(E), vp(V), vp(scal),
                                                   vp(post), vp(info))
            _check(rc)
        scalars = {nm: scal[:, i].copy() for i, nm in enumerate(names)}
        return SweepResult(E, V, scalars, post, info, V_dev=[(r0, r0 + n_pts, V_dev)] if to_dev else None)


def visible_devices():
    import torch
    return list(range(torch.cuda.device_count())) if torch.cuda.is_available() else []


class EnginePool:
    """One plan on several GPUs of this box, driven from ONE process: the sweep driver of the reference
    (numerical_system.py:885-928) shards the flattened grid into contiguous blocks, one host thread per GPU calls
    ``scqed_sweep_run_host`` on its block (ctypes releases the GIL) and every block lands in its rows of the same
    page-locked result arrays.  Sweep points are independent: no collective, no peer traffic."""

    def __init__(self, plan, devices=None):
        devs = visible_devices() if devices is None else [int(d) for d in devices]
        if not devs:
            raise EngineError("no CUDA device visible; the sweep engine has no CPU fallback")
        self.plan = plan
        self.devices = devs
        self.engines = [Engine(plan, device=d) for d in devs]
        self.n = self.engines[0].n

    def close(self):
        for e in self.engines:
            e.close()
        self.engines = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sweep(self, params, k, want_vectors=False, resonator=False, min_points_per_gpu=256, **kw) -> SweepResult:
        from .distributed import shard_bounds
        params = np.ascontiguousarray(params, dtype=np.float64)
        n_pts = params.shape[0]
        use = max(1, min(len(self.engines), n_pts // max(1, int(min_points_per_gpu))))
        if use == 1 or kw.get("matrix_elements"):
            return self.engines[0].sweep(params, k, want_vectors=want_vectors, resonator=resonator, **kw)
        eng0 = self.engines[0]
        to_dev = bool(kw.get("vectors_on_device")) and want_vectors
        real_vectors = kw.pop("real_vectors", False)
        ns = self.plan.program.n_scalar
        out = {"E": pinned_empty((n_pts, k), np.float64), "scalars": pinned_empty((n_pts, max(ns, 1)), np.float64),
               "info": pinned_empty((n_pts,), np.int32)}
        if resonator:
            out["Erwa"] = pinned_empty((n_pts, int(self.plan.post["resonator"]["nmax"]) + 1 + k, k), np.float64)

        def run(real):
            import threading
            if want_vectors and not to_dev:
                out["V"] = pinned_empty((n_pts, k, self.n), np.float64 if real else np.complex128)
            res, errs = [None] * use, []

            def work(i):
                lo, hi = shard_bounds(n_pts, use, i)
                if hi <= lo:
                    return
                try:
                    res[i] = self.engines[i].sweep(params[lo:hi], k, want_vectors=want_vectors, resonator=resonator,
                                                   out=out, row_offset=lo, real_vectors=bool(real), **kw)
                except Exception as exc:            # re-raised on the calling thread
                    errs.append(exc)
            th = [threading.Thread(target=work, args=(i,)) for i in range(use)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            return res, errs

        real_req = bool(real_vectors) and want_vectors and not eng0.rotated and not to_dev
        res, errs = run(real_req)
        if errs and real_req and real_vectors == "auto" and all("eigenvectors are complex" in str(e) for e in errs):
            res, errs = run(False)
        if errs:
            raise errs[0]
        names = list(self.plan.program.scalar_names)
        scalars = {nm: out["scalars"][:, i].copy() for i, nm in enumerate(names)}
        V_dev = [r.V_dev[0] for r in res if r is not None and r.V_dev] if to_dev else None
        return SweepResult(out["E"], out.get("V"), scalars, out.get("Erwa"), out["info"], V_dev=V_dev)


def pauli_coefficients(E0, E1, Op, device=None):
    """``util.pauliCoefficients`` (util.py:300-391) on the device: ``E0``, ``E1`` [n_pts] and the 2x2
    computational-basis operators ``Op`` [n_pts, 2, 2] (complex) -> (hx, hy, hz) numpy arrays."""
    import torch
    lib = load_library()
    dev = _require_cuda(device)
    e01 = torch.as_tensor(np.ascontiguousarray(np.stack([np.asarray(E0, dtype=np.float64), np.asarray(E1, dtype=np.float64)], axis=1)), device=dev)
    op = torch.as_tensor(np.ascontiguousarray(np.asarray(Op, dtype=np.complex128).reshape(-1, 4)), device=dev)
    n_pts = e01.shape[0]
    if op.shape[0] != n_pts:
        raise EngineError("Op must hold one 2x2 operator per point (%d), got %d" % (n_pts, op.shape[0]))
    h = torch.empty((n_pts, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _check(lib.scqed_pauli_coefficients(dev.index, e01.data_ptr(), op.data_ptr(), n_pts, h.data_ptr(), _stream_ptr(dev)))
    h = h.cpu().numpy()
    return h[:, 0].copy(), h[:, 1].copy(), h[:, 2].copy()


def dense_matrix_elements(Op, W, device=None):
    """``<W[p, a]| Op |W[p, b]>`` for one fixed dense operator ``Op`` [n, n] and ``W`` [n_pts, m, n] (m <= 4 vectors
    per point): complex128 numpy array [n_pts, m, m] (``scqed_dense_matrix_elements``)."""
    import torch
    lib = load_library()
    dev = _require_cuda(device)
    A = torch.as_tensor(np.ascontiguousarray(np.asarray(Op, dtype=np.complex128)), device=dev)
    Wd = W if isinstance(W, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(np.asarray(W, dtype=np.complex128)), device=dev)
    Wd = Wd.to(device=dev, dtype=torch.complex128).contiguous()
    if A.ndim != 2 or A.shape[0] != A.shape[1] or Wd.ndim != 3 or Wd.shape[2] != A.shape[0]:
        raise EngineError("Op must be [n, n] and W [n_pts, m, n]")
    n_pts, m, n = Wd.shape
    out = torch.empty((n_pts, m, m), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        _check(lib.scqed_dense_matrix_elements(dev.index, A.data_ptr(), Wd.data_ptr(), n, m, n_pts, out.data_ptr(), _stream_ptr(dev)))
    return out.cpu().numpy()


def eigh_batched(H, k, want_vectors=False, solver=SOLVER_AUTO):
    """Lowest-k eigenpairs of a batch of Hermitian matrices on the device
    (``diagonalize`` / ``util.diagDenseH`` parity hook).  ``H``: torch complex128 [B, n, n] on a
    CUDA device; it is overwritten.  Returns (E [B,k], V [B,k,n] | None, info [B])."""
    import torch
    lib = load_library()
    if not (isinstance(H, torch.Tensor) and H.is_cuda and H.dtype == torch.complex128 and H.ndim == 3):
        raise EngineError("H must be a CUDA complex128 tensor [batch, n, n]")
    H = H.contiguous()
    B, n, _ = H.shape
    E = torch.empty((B, k), dtype=torch.float64, device=H.device)
    V = torch.empty((B, k, n), dtype=torch.complex128, device=H.device) if want_vectors else None
    info = torch.empty((B,), dtype=torch.int32, device=H.device)
    with torch.cuda.device(H.device):
        _check(lib.scqed_eigh_batched(H.device.index, H.data_ptr(), n, B, k, E.data_ptr(),
                                      V.data_ptr() if V is not None else None, info.data_ptr(), int(solver), _stream_ptr(H.device)))
    return E, V, info
