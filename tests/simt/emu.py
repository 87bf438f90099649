"""TEST INFRASTRUCTURE: ctypes binding of tests/simt/libdre_emu.so (the product's CUDA kernel sources compiled
for the host-side SIMT emulator).  Everything here works in the SOLVER ordering of the symbolic analysis, as the
device code does; `perm[new] = old` converts."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

from . import build_emu

_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is None:
        # DRE_EMU_LIB: an alternative build of the same sources (e.g. with -fsanitize=address, see tests/simt/README.md)
        _lib = C.CDLL(os.environ.get("DRE_EMU_LIB") or build_emu.build())
        _lib.emu_open.argtypes = [C.c_int64, _lp, _lp, _dp, _lp, _lp, _dp, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_void_p)]
        _lib.emu_close.argtypes = [C.c_void_p]
        _lib.emu_sizes.argtypes = [C.c_void_p, _lp]
        _lib.emu_perm.argtypes = [C.c_void_p, _ip]
        _lib.emu_factor.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
        _lib.emu_get.argtypes = [C.c_void_p, C.c_int, _dp]
        _lib.emu_sweeps.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int, _dp, C.c_int64, C.c_int, _dp, C.c_int64]
        _lib.emu_smw.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_double, _dp, C.c_int64, C.c_int, C.c_double,
                                 _dp, C.c_int64, _dp, C.c_int64, C.POINTER(C.c_int)]
        _lib.emu_spmm.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp, C.c_int64, C.c_double, _dp, C.c_int64, C.c_int]
        _lib.emu_gram.argtypes = [_dp, C.c_int64, C.c_int, _dp, C.c_int64, C.c_int, C.c_int64, _dp, _dp, C.c_int64, _dp,
                                  C.c_int64, C.c_int]
        _lib.emu_tall_gemm.argtypes = [C.c_double, _dp, C.c_int64, C.c_int, _dp, C.c_int64, C.c_int, C.c_double, _dp,
                                       C.c_int64, C.c_int, C.c_int64]
        _lib.emu_pivchol.argtypes = [_dp, C.c_int64, C.c_int, C.c_double, C.c_double, _dp, _ip, _dp]
        _lib.emu_norm_diag.argtypes = [_dp, C.c_int64, C.c_int, _dp, _dp]
        _lib.emu_colnorm2.argtypes = [_dp, C.c_int64, C.c_int64, C.c_int, C.c_int, _dp]
        _lib.emu_panel_transposes.argtypes = [_dp, C.c_int64, _dp, C.c_int64, C.c_int64, C.c_int, _ip, _dp, C.c_int64]
        _lib.emu_counters.argtypes = [C.POINTER(C.c_long)]
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _rm(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Solver:
    """Symbolic analysis + emulated numeric factorization / sweeps of  a*A + emu*E."""

    def __init__(self, E, A, leaf=0, maxsn=0):
        E, A = sp.csc_matrix(E), sp.csc_matrix(A)
        E.sort_indices()
        A.sort_indices()
        self.n = E.shape[0]
        keep = [E.indptr.astype(np.int64), E.indices.astype(np.int64), E.data.astype(np.float64),
                A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)]
        self.h = C.c_void_p()
        rc = lib().emu_open(self.n, keep[0].ctypes.data_as(_lp), keep[1].ctypes.data_as(_lp), _d(keep[2]),
                            keep[3].ctypes.data_as(_lp), keep[4].ctypes.data_as(_lp), _d(keep[5]), 0, leaf, maxsn,
                            C.byref(self.h))
        if rc:
            raise RuntimeError("emu_open failed")
        sizes = np.zeros(8, dtype=np.int64)
        lib().emu_sizes(self.h, sizes.ctypes.data_as(_lp))
        self.sizes = dict(zip(["n", "nsn", "nlevels", "nnz_L", "linv_elems", "upd_elems", "sum_u", "max_sn"],
                              map(int, sizes)))
        self.perm = np.zeros(self.n, dtype=np.int32)
        lib().emu_perm(self.h, self.perm.ctypes.data_as(_ip))
        self.cplx = False

    def close(self):
        if self.h:
            lib().emu_close(self.h)
            self.h = None

    def factor(self, a, emu, m21=False):
        """m21: leave M21 = L21 Linv in the panels; sweeps() then runs the row-split kernels (sweep v2)."""
        self.cplx = isinstance(emu, complex)
        return lib().emu_factor(self.h, int(self.cplx), float(a), float(np.real(emu)), float(np.imag(emu)), int(m21))

    def get(self, what):
        cnt = {"L": self.sizes["nnz_L"], "Linv": self.sizes["linv_elems"], "dvec": self.n}[what]
        buf = np.zeros(cnt * (2 if self.cplx else 1))
        lib().emu_get(self.h, {"L": 0, "Linv": 1, "dvec": 2}[what], _d(buf))
        return buf.view(np.complex128) if self.cplx else buf

    def sweeps(self, R, Vt=None):
        """R, Vt: real blocks in ORIGINAL row order; returns the solved block [R, Vt] in original order."""
        Rp = _rm(R[self.perm])
        r = Rp.shape[1]
        m = 0 if Vt is None else Vt.shape[1]
        Vp = _rm(Vt[self.perm]) if m else None
        ldw = (r + m + 3) & ~3
        tw = 2 if self.cplx else 1
        W = np.zeros((self.n, ldw * tw))
        lib().emu_sweeps(self.h, _d(Rp), Rp.shape[1], r, _d(Vp), m, m, _d(W), ldw)
        Wv = W.view(np.complex128) if self.cplx else W
        out = np.empty((self.n, r + m), dtype=Wv.dtype)
        out[self.perm] = Wv[:, :r + m]
        return out

    def spmm(self, which, alpha, X, beta, Y):
        Xp, Yp = _rm(X[self.perm]), _rm(Y[self.perm])
        lib().emu_spmm(self.h, {"A": 0, "E": 1}[which], alpha, _d(Xp), Xp.shape[1], beta, _d(Yp), Yp.shape[1],
                       Xp.shape[1])
        out = np.empty_like(Yp)
        out[self.perm] = Yp
        return out


def gram(X, Y, roww=None, sm_count=148):
    X, Y = _rm(X), _rm(Y)
    out = np.full((X.shape[1], Y.shape[1]), np.nan)
    w = _rm(roww) if roww is not None else None
    lib().emu_gram(_d(X), X.shape[1], X.shape[1], _d(Y), Y.shape[1], Y.shape[1], X.shape[0], _d(w), _d(out),
                   out.shape[1], None, 0, sm_count)
    return out


def tall_gemm(alpha, X, W, beta, Y, w_trans=False):
    X, W, Y = _rm(X), _rm(W), _rm(Y).copy()
    a, b = X.shape[1], Y.shape[1]
    lib().emu_tall_gemm(alpha, _d(X), a, a, _d(W), W.shape[1], int(w_trans), beta, _d(Y), b, b, X.shape[0])
    return Y


def pivchol(G, drop2, rel2):
    """k_pivchol on a pb x pb Gram matrix (pb <= 64): returns (nsel, Wsel (pb x 64), dfirst, remaining)."""
    G = _rm(G)
    pb = G.shape[0]
    W = np.full((64, 64), np.nan)      # the kernel owns a full 64 x 64 block (context.cu: wsel.ensure(PB * PB))
    info = np.zeros(4, dtype=np.int32)
    dinfo = np.zeros(4)
    lib().emu_pivchol(_d(G), pb, pb, float(drop2), float(rel2), _d(W), info.ctypes.data_as(_ip), _d(dinfo))
    return int(info[0]), W[:pb], float(dinfo[0]), float(dinfo[1])


def norm_diag(G, t):
    G, t = _rm(G), _rm(t)
    out = np.zeros(1)
    lib().emu_norm_diag(_d(G), G.shape[1], G.shape[0], _d(t), _d(out))
    return float(out[0])


def colnorm2(P, nblk=7):
    P = _rm(P)
    out = np.full(P.shape[1], np.nan)
    lib().emu_colnorm2(_d(P), P.shape[1], P.shape[0], P.shape[1], nblk, _d(out))
    return out


def panel_roundtrip(M, iperm):
    """column-major host matrix -> row-major permuted panel -> back (the upload / download kernels)."""
    M = np.asfortranarray(M, dtype=np.float64)
    n, cols = M.shape
    ld = (cols + 1) & ~1
    panel = np.full((n, ld), np.nan)
    back = np.full((n, cols), np.nan, order="F")
    ip = np.ascontiguousarray(iperm, dtype=np.int32)
    lib().emu_panel_transposes(_d(panel), ld, _d(M), n, n, cols, ip.ctypes.data_as(_ip), _d(back), n)
    return panel[:, :cols], back


def set_diag_narrow_min(v):
    lib().emu_set_diag_narrow_min(int(v))


def set_diag_variant(v):
    lib().emu_set_diag_variant(int(v))


def set_spmm_variant(v):
    lib().emu_set_spmm_variant(int(v))


def counters():
    c = (C.c_long * 3)()
    lib().emu_counters(c)
    return {"launches": c[0], "ctas": c[1], "switches": c[2]}
