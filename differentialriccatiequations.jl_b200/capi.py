"""ctypes binding of the C ABI declared in include/dre_b200.h (libdre_b200.so).

This is the only door from the Python host mirror into CUDA -- the analogue of the ``ccall``
stubs of the Julia glue shown in INTEGRATION.md.  There is no CPU fallback: if the library is
missing or no CUDA device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdre_b200.so")

EXPORTED_SYMBOLS = [
    "dre_last_error", "dre_version", "dre_symbolic_create", "dre_symbolic_destroy", "dre_symbolic_get_info",
    "dre_symbolic_export", "dre_create", "dre_destroy", "dre_sync", "dre_set_pencil", "dre_get_symbolic_info",
    "dre_mat_create", "dre_mat_free", "dre_mat_upload", "dre_mat_download", "dre_mat_copy", "dre_mat_axpby",
    "dre_spmm", "dre_gemm_tn", "dre_gemm_nn", "dre_set_operator", "dre_prefactor", "dre_shift_solve", "dre_adi_step", "dre_adi_solve", "dre_mat_devptr", "dre_get_stream", "dre_set_dense_only", "dre_mat_wrap",
    "dre_ldlt_norm", "dre_ldlt_norm_begin", "dre_ldlt_norm_end", "dre_ldlt_compress", "dre_compress_begin", "dre_compress_scale_hint", "dre_compress_add", "dre_compress_finish", "dre_hint_orthonormal", "dre_rrqr", "dre_debug_export", "dre_debug_eigh", "dre_timer_start", "dre_timer_stop", "dre_stats_reset",
    "dre_stats_get", "dre_arnoldi_orth",
]


class DreError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdre_b200 error {code}: {msg}")
        self.code = code


class View(C.Structure):
    _fields_ = [("id", C.c_int32), ("col0", C.c_int32), ("ncols", C.c_int32)]


class SymbolicInfo(C.Structure):
    _fields_ = [("n", C.c_int64), ("nnz_pattern", C.c_int64), ("nnz_L", C.c_int64),
                ("sum_update_rows", C.c_int64), ("factor_flops", C.c_double), ("nsupernodes", C.c_int32),
                ("nlevels", C.c_int32), ("max_front", C.c_int32), ("max_supernode", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("factorizations", C.c_int64), ("solves", C.c_int64),
                ("spmms", C.c_int64), ("grams", C.c_int64), ("tallgemms", C.c_int64),
                ("prefactors", C.c_int64), ("prefactor_hits", C.c_int64),
                ("ms_factor", C.c_double), ("ms_solve", C.c_double), ("ms_spmm", C.c_double),
                ("ms_gram", C.c_double), ("ms_tallgemm", C.c_double),
                ("flops_factor", C.c_double), ("flops_gram", C.c_double), ("flops_tallgemm", C.c_double),
                ("flops_solve", C.c_double), ("bytes_solve", C.c_double), ("bytes_spmm", C.c_double),
                ("bytes_gram", C.c_double), ("bytes_tallgemm", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load():
    """Load libdre_b200.so (built in-tree by build.py).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DreError(-100, f"{LIB_PATH} not found: run `python __graft_entry__.py build` "
                             "(there is no CPU fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    p = C.c_void_p
    i32, i64, dbl = C.c_int32, C.c_int64, C.c_double
    pi64 = C.POINTER(C.c_int64)
    pdbl = C.POINTER(C.c_double)
    lib.dre_last_error.restype = C.c_char_p
    lib.dre_last_error.argtypes = [p]
    lib.dre_version.restype = C.c_char_p
    lib.dre_symbolic_create.argtypes = [i64, pi64, pi64, pdbl, pi64, pi64, pdbl, i32, i32, C.POINTER(p)]
    lib.dre_symbolic_destroy.argtypes = [p]
    lib.dre_symbolic_destroy.restype = None
    lib.dre_symbolic_get_info.argtypes = [p, C.POINTER(SymbolicInfo)]
    lib.dre_symbolic_export.argtypes = [p, C.c_char_p, p, i64, pi64]
    lib.dre_create.argtypes = [i32, C.POINTER(p)]
    lib.dre_destroy.argtypes = [p]
    lib.dre_sync.argtypes = [p]
    lib.dre_set_pencil.argtypes = [p, i64, pi64, pi64, pdbl, pi64, pi64, pdbl, i32]
    lib.dre_get_symbolic_info.argtypes = [p, C.POINTER(SymbolicInfo)]
    lib.dre_mat_create.argtypes = [p, i32, C.POINTER(i32)]
    lib.dre_mat_free.argtypes = [p, i32]
    lib.dre_mat_upload.argtypes = [p, View, pdbl, i64]
    lib.dre_mat_download.argtypes = [p, View, pdbl, i64]
    lib.dre_mat_copy.argtypes = [p, View, View]
    lib.dre_mat_axpby.argtypes = [p, dbl, View, dbl, View]
    lib.dre_arnoldi_orth.argtypes = [p, View, View, View, pdbl]
    lib.dre_spmm.argtypes = [p, i32, dbl, View, dbl, View]
    lib.dre_gemm_tn.argtypes = [p, View, View, pdbl, i64]
    lib.dre_gemm_nn.argtypes = [p, dbl, View, pdbl, i64, dbl, View]
    lib.dre_set_operator.argtypes = [p, dbl, dbl, dbl, View, View]
    lib.dre_prefactor.argtypes = [p, dbl, dbl]
    lib.dre_shift_solve.argtypes = [p, dbl, dbl, View, View, View]
    lib.dre_adi_step.argtypes = [p, dbl, dbl, View, View, View]
    lib.dre_adi_solve.argtypes = [p, dbl, dbl, View, View, View]
    lib.dre_mat_devptr.argtypes = [p, View, C.POINTER(p), pi64]
    lib.dre_get_stream.argtypes = [p, C.POINTER(p)]
    lib.dre_set_dense_only.argtypes = [p, i64]
    lib.dre_mat_wrap.argtypes = [p, p, i64, i32, C.POINTER(i32)]
    lib.dre_ldlt_norm.argtypes = [p, View, pdbl, i64, dbl, pdbl]
    lib.dre_ldlt_norm_begin.argtypes = [p, View, pdbl, dbl]
    lib.dre_ldlt_norm_end.argtypes = [p, pdbl]
    lib.dre_ldlt_compress.argtypes = [p, i32, C.POINTER(View), C.POINTER(pdbl), pi64, pdbl, dbl, View, pdbl,
                                      C.POINTER(i32)]
    lib.dre_compress_begin.argtypes = [p, i32, dbl]
    lib.dre_compress_scale_hint.argtypes = [p, dbl]
    lib.dre_compress_add.argtypes = [p, i32, C.POINTER(View), C.POINTER(pdbl), pi64, pdbl]
    lib.dre_compress_finish.argtypes = [p, View, pdbl, C.POINTER(i32)]
    lib.dre_hint_orthonormal.argtypes = [p, View]
    lib.dre_rrqr.argtypes = [p, i32, C.POINTER(View), dbl, dbl, View, pdbl, i64, C.POINTER(i32)]
    lib.dre_debug_export.argtypes = [p, C.c_char_p, p, i64, pi64]
    lib.dre_debug_eigh.argtypes = [p, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.dre_timer_start.argtypes = [p]
    lib.dre_timer_stop.argtypes = [p, pdbl]
    lib.dre_stats_reset.argtypes = [p, i32]
    lib.dre_stats_get.argtypes = [p, C.POINTER(Stats)]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("dre_last_error", "dre_version", "dre_symbolic_destroy"):
            fn.restype = i32
    _lib = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _csc_arrays(M):
    M = M.tocsc()
    M.sort_indices()
    return (np.ascontiguousarray(M.indptr, dtype=np.int64), np.ascontiguousarray(M.indices, dtype=np.int64),
            np.ascontiguousarray(M.data, dtype=np.float64))


def _check(lib, ctx, rc):
    if rc != 0:
        msg = lib.dre_last_error(ctx)
        raise DreError(rc, msg.decode() if msg else "?")


class SymbolicAnalysis:
    """Host-only symbolic analysis (no GPU needed): dre_symbolic_* functions."""

    def __init__(self, E, A, leaf_size=0):
        self.lib = load()
        ep, ei, ev = _csc_arrays(E)
        ap, ai, av = _csc_arrays(A)
        n = E.shape[0]
        self.h = C.c_void_p()
        rc = self.lib.dre_symbolic_create(n, _iptr(ep), _iptr(ei), _dptr(ev), _iptr(ap), _iptr(ai), _dptr(av), 0,
                                          leaf_size, C.byref(self.h))
        _check(self.lib, None, rc)

    def info(self):
        info = SymbolicInfo()
        _check(self.lib, None, self.lib.dre_symbolic_get_info(self.h, C.byref(info)))
        return info.as_dict()

    def export(self, name):
        ln = C.c_int64(0)
        _check(self.lib, None, self.lib.dre_symbolic_export(self.h, name.encode(), None, 0, C.byref(ln)))
        isf = name in ("asm_a", "asm_e", "csr_a", "csr_e")
        buf = np.empty(ln.value, dtype=np.float64 if isf else np.int64)
        _check(self.lib, None,
               self.lib.dre_symbolic_export(self.h, name.encode(), buf.ctypes.data_as(C.c_void_p), ln.value,
                                            C.byref(ln)))
        return buf

    def __del__(self):
        try:
            if self.h:
                self.lib.dre_symbolic_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Context:
    """One GPU, one stream (dre_create / dre_destroy)."""

    def __init__(self, device=0):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.dre_create(device, C.byref(self.h))
        if rc != 0:
            msg = self.lib.dre_last_error(None)
            raise DreError(rc, msg.decode() if msg else "?")
        self.n = None

    def check(self, rc):
        _check(self.lib, self.h, rc)

    def close(self):
        if self.h:
            self.lib.dre_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self.check(self.lib.dre_sync(self.h))

    def set_pencil(self, E, A):
        ep, ei, ev = _csc_arrays(E)
        ap, ai, av = _csc_arrays(A)
        n = E.shape[0]
        self.check(self.lib.dre_set_pencil(self.h, n, _iptr(ep), _iptr(ei), _dptr(ev), _iptr(ap), _iptr(ai),
                                           _dptr(av), 0))
        self.n = n

    def symbolic_info(self):
        info = SymbolicInfo()
        self.check(self.lib.dre_get_symbolic_info(self.h, C.byref(info)))
        return info.as_dict()

    def debug_export(self, what, complex_valued=False):
        """Raw numeric-factorization arrays ("L", "Linv", "dvec", "U") of the factorization currently held."""
        ln = C.c_int64(0)
        self.check(self.lib.dre_debug_export(self.h, what.encode(), None, 0, C.byref(ln)))
        buf = np.empty(ln.value // (16 if complex_valued else 8), dtype=np.complex128 if complex_valued else np.float64)
        self.check(self.lib.dre_debug_export(self.h, what.encode(), buf.ctypes.data_as(C.c_void_p), ln.value,
                                             C.byref(ln)))
        return buf

    def debug_eigh(self, A):
        """The library's own symmetric eigensolver on a host matrix: (ascending eigenvalues, eigenvector columns)."""
        A = np.asfortranarray(np.asarray(A, dtype=np.float64))
        k = A.shape[0]
        w = np.empty(k)
        V = np.empty((k, k), order="F")
        self.check(self.lib.dre_debug_eigh(self.h, k, _dptr(A), _dptr(w), _dptr(V)))
        return w, V

    def timer_start(self):
        self.check(self.lib.dre_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double(0.0)
        self.check(self.lib.dre_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def stats_reset(self, timing=False):
        self.check(self.lib.dre_stats_reset(self.h, 1 if timing else 0))

    def stats(self):
        s = Stats()
        self.check(self.lib.dre_stats_get(self.h, C.byref(s)))
        return s.as_dict()
