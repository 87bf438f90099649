"""Host-side mirror of the reference's Julia API for the LRSIF-ADI hot path, over the C ABI.

Names, argument meaning, control flow, observer call order and error behaviour follow the reference
(paths relative to /root/reference/; Julia's ``f!`` is spelled ``f_`` here):

  lowrank / LDLt / concatenate_ / compress_ / norm           src/LDLt.jl
  LowRankUpdate / lr_update                                  src/LowRankUpdate.jl
  BlockLinearProblem / Backslash / ShermanMorrisonWoodbury   src/blocklinear/*.jl
  Shifts.{Projection, Heuristic, Cyclic, Wrapped, ...}       src/Shifts.jl, src/shifts/*.jl
  GALEProblem / ADI / init / step_ / solve_ / solve          src/lyapunov/{types,adi,residual}.jl
  GDREProblem / Ros1 / Ros2 / GAREProblem / Newton / solve   src/riccati/*.jl
  observers                                                  src/Callbacks.jl (method names without "!")

Outer factors (the n x k matrices) live on the GPU as ``DeviceMatrix`` panels; the small k x k cores
live on the host as NumPy arrays -- the "xpu" layout of the reference's own GPU test
(test/cuda.jl:66, GPU outer factor + CPU core).  All heavy arithmetic goes through
``libdre_b200.so``; there is no CPU fallback.  Tiny dense eigen/SVD problems of the Projection
shift strategy run on the host LAPACK exactly like the reference does
(src/shifts/projection.jl:63-67 "Ensure data lives on the CPU").
"""
from __future__ import annotations

import ctypes as C
import itertools
import math
import os as _os
import warnings

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from . import capi
from . import dist as _dist
from .capi import View

EPS = float(np.finfo(np.float64).eps)

# =============================================================================================
# backend: one context per process, device panels
# =============================================================================================
_backend = None


# DRE_CHECK_PENCIL=0 skips the content check when the very same matrix objects are passed again
_CHECK_PENCIL_VALUES = _os.environ.get("DRE_CHECK_PENCIL", "1") not in ("", "0")


class Backend:
    def __init__(self, device=0):
        self.device = device
        self.ctx = capi.Context(device)
        self.lib = self.ctx.lib
        self.h = self.ctx.h
        self.generation = 0
        self.pencil_key = None
        self.E = self.A = None
        self.n = None

    @staticmethod
    def _pencil_key(E, A):
        """Content key of the pencil (shape, nnz and a CRC of structure and values): an equal copy of the resident
        pencil is recognised (no re-upload that would invalidate every DeviceMatrix), values changed in place are too
        (no stale factorization).  ~4 ms at n = 79 841; evaluated once per solve / init, never per ADI step."""
        import zlib

        def crc(M):
            c = 0
            for arr in (M.indptr, M.indices, M.data):
                c = zlib.crc32(np.ascontiguousarray(arr).view(np.uint8), c)
            return c

        return (E.shape, A.shape, E.nnz, A.nnz, E.format, A.format, crc(E), crc(A))

    def ensure_pencil(self, E, A):
        if not (sp.issparse(E) and sp.issparse(A)):
            raise TypeError("E and A must be scipy sparse matrices")
        def quick(M):   # cheap signature for the hot case below (0.3 ms at n = 79 841)
            return (M.nnz, float(M.data.sum()), float(np.dot(M.data, M.data)))

        if getattr(self, "E", None) is E and getattr(self, "A", None) is A and self.pencil_key is not None:
            # the very objects of the last call (every time step of a Rosenbrock solve): only an in-place change of
            # their values is possible, which the sums catch
            if not _CHECK_PENCIL_VALUES or self._pencil_quick == (quick(E), quick(A)):
                return
        key = self._pencil_key(E, A)
        if self.pencil_key == key:
            self.E, self.A, self._pencil_quick = E, A, (quick(E), quick(A))
            return
        self.ctx.set_pencil(E, A)
        self.generation += 1
        self.pencil_key = key
        self.E, self.A, self._pencil_quick = E, A, (quick(E), quick(A))
        self.n = E.shape[0]

    def check(self, rc):
        self.ctx.check(rc)

    def scratch(self, cols: int) -> "DeviceMatrix":
        """View of a persistent scratch panel with at least `cols` columns (valid until the next call)."""
        sp_ = getattr(self, "_scratch", None)
        if sp_ is None or sp_.gen != self.generation or sp_.cols < cols:
            self._scratch = None  # release the old one first
            sp_ = self._scratch = _Panel(self, max(cols + cols // 2, 512))
        return DeviceMatrix(sp_, 0, cols)


class _LaneBackend(Backend):
    """Second context on the same device for the dense low-rank algebra only (dre_set_dense_only): the compression
    lane of DRE_ASYNC_COMPRESS.  Its panels either belong to its own arena or wrap memory of the main context."""

    def __init__(self, main: Backend):
        super().__init__(main.device)
        self.n = main.n
        self.generation = 1
        self.check(self.lib.dre_set_dense_only(self.h, main.n))


class _PipeLaneBackend(Backend):
    """Rank 1 of the multi-GPU pipeline mode (dre_b200.dist): a dense-only context that holds X and runs compress!."""

    def __init__(self, n: int, device):
        super().__init__(0 if device is None else device)
        self.n = n
        self.generation = 1
        self.check(self.lib.dre_set_dense_only(self.h, n))


def _pipe_lane_backend(n: int, device) -> Backend:
    return _PipeLaneBackend(n, device)


def _lane(main: Backend) -> _LaneBackend:
    lane = getattr(main, "_lane_be", None)
    if lane is None or lane.n != main.n or getattr(main, "_lane_gen", None) != main.generation:
        if lane is not None:
            lane.ctx.close()
        lane = main._lane_be = _LaneBackend(main)
        main._lane_gen = main.generation
    return lane


def backend(device=None) -> Backend:
    global _backend
    if _backend is None:
        _backend = Backend(0 if device is None else device)
    return _backend


def reset_backend():
    global _backend
    if _backend is not None:
        lane = getattr(_backend, "_lane_be", None)
        if lane is not None:
            lane.ctx.close()
        _backend.ctx.close()
    _backend = None


class _Panel:
    """Owner of one device panel (freed when the last view dies)."""

    def __init__(self, be: Backend, cols: int):
        self.be, self.cols, self.gen = be, cols, be.generation
        self.version = 0   # bumped by every in-place writer (_touch): validates the "orthonormal columns" hint
        pid = C.c_int32(-1)
        be.check(be.lib.dre_mat_create(be.h, cols, C.byref(pid)))
        self.id = pid.value

    def __del__(self):
        try:
            if self.be.ctx.h and self.gen == self.be.generation:
                self.be.lib.dre_mat_free(self.be.h, self.id)
        except Exception:
            pass


class _WrappedPanel:
    """Panel of context `be` that aliases the memory of a DeviceMatrix owned by another context (dre_mat_wrap).
    Keeps the owner alive; synchronises `be` before letting go, so that the owner's arena cannot hand the memory out
    again while work queued through the alias is still running."""

    def __init__(self, be: Backend, src: "DeviceMatrix"):
        self.be, self.cols, self.gen, self.src = be, src.ncols, be.generation, src
        self.version = 0
        owner = src.panel.be
        ptr, ld = C.c_void_p(), C.c_int64()
        owner.check(owner.lib.dre_mat_devptr(owner.h, src.view, C.byref(ptr), C.byref(ld)))
        pid = C.c_int32(-1)
        be.check(be.lib.dre_mat_wrap(be.h, ptr, ld.value, src.ncols, C.byref(pid)))
        self.id = pid.value

    def __del__(self):
        try:
            if self.be.ctx.h and self.gen == self.be.generation:
                self.be.lib.dre_sync(self.be.h)
                self.be.lib.dre_mat_free(self.be.h, self.id)
        except Exception:
            pass


def _alias(be: Backend, src: "DeviceMatrix") -> "DeviceMatrix":
    return DeviceMatrix(_WrappedPanel(be, src), 0, src.ncols)


class DeviceMatrix:
    """n x k real matrix on the GPU: a column range of a panel (cf. dre_view)."""

    def __init__(self, panel: _Panel, col0: int, ncols: int):
        self.panel, self.col0, self.ncols = panel, col0, ncols

    @staticmethod
    def empty(cols: int) -> "DeviceMatrix":
        be = backend()
        return DeviceMatrix(_Panel(be, max(cols, 0)), 0, cols)

    @staticmethod
    def from_host(M) -> "DeviceMatrix":
        M = np.asfortranarray(np.asarray(M, dtype=np.float64))
        if M.ndim == 1:
            M = M.reshape(-1, 1, order="F")
        be = backend()
        if M.shape[0] != be.n:
            raise ValueError(f"row count {M.shape[0]} does not match the pencil dimension {be.n}")
        d = DeviceMatrix.empty(M.shape[1])
        if M.shape[1]:
            be.check(be.lib.dre_mat_upload(be.h, d.view, capi._dptr(M), M.shape[0]))
        return d

    @property
    def view(self) -> View:
        if self.panel.gen != self.panel.be.generation:
            raise RuntimeError("stale DeviceMatrix: the pencil of the backend was replaced")
        return View(self.panel.id, self.col0, self.ncols)

    @property
    def shape(self):
        return (self.panel.be.n, self.ncols)

    def cols(self, a: int, b: int) -> "DeviceMatrix":
        assert 0 <= a <= b <= self.ncols
        return DeviceMatrix(self.panel, self.col0 + a, b - a)

    def to_host(self) -> np.ndarray:
        be = self.panel.be
        out = np.empty((be.n, self.ncols), dtype=np.float64, order="F")
        if self.ncols:
            be.check(be.lib.dre_mat_download(be.h, self.view, capi._dptr(out), be.n))
        return out

    def copy(self) -> "DeviceMatrix":
        be = self.panel.be
        d = DeviceMatrix(_Panel(be, max(self.ncols, 0)), 0, self.ncols)
        if self.ncols:
            be.check(be.lib.dre_mat_copy(be.h, d.view, self.view))
        return d

    def __array__(self, dtype=None, copy=None):
        return self.to_host()


def _touch(M: DeviceMatrix):
    """Record an in-place write into M's panel: an "orthonormal columns" mark taken before it no longer holds
    (the residual factor a compress! returned is overwritten by every ADI step, adi.jl:171)."""
    M.panel.version += 1


def _mark_orthonormal(M: DeviceMatrix):
    M._ortho_version = M.panel.version


def _is_orthonormal(M) -> bool:
    return getattr(M, "_ortho_version", None) == M.panel.version


def _as_device(M) -> DeviceMatrix:
    return M if isinstance(M, DeviceMatrix) else DeviceMatrix.from_host(M)


def _ncols(L) -> int:
    return L.ncols if isinstance(L, DeviceMatrix) else int(np.shape(L)[1])


def hcat(Xs) -> DeviceMatrix:
    """src/util/_hcat.jl:5-18 on device."""
    Xs = list(Xs)
    be = backend()
    out = DeviceMatrix.empty(sum(X.ncols for X in Xs))
    k = 0
    for X in Xs:
        if X.ncols:
            be.check(be.lib.dre_mat_copy(be.h, out.cols(k, k + X.ncols).view, X.view))
        k += X.ncols
    return out


def spmm(op: str, X: DeviceMatrix, alpha=1.0, Y: DeviceMatrix | None = None, beta=0.0) -> DeviceMatrix:
    """Y = alpha * op * X + beta * Y, op in {'E','A'} (symmetric pencil: E'X == EX)."""
    be = backend()
    if Y is None:
        Y = DeviceMatrix.empty(X.ncols)
        beta = 0.0
    be.check(be.lib.dre_spmm(be.h, ord(op), float(alpha), X.view, float(beta), Y.view))
    _touch(Y)
    return Y


def gemm_tn(X: DeviceMatrix, Y: DeviceMatrix) -> np.ndarray:
    """X' * Y -> host matrix."""
    be = backend()
    out = np.zeros((X.ncols, Y.ncols), dtype=np.float64, order="F")
    if X.ncols and Y.ncols:
        be.check(be.lib.dre_gemm_tn(be.h, X.view, Y.view, capi._dptr(out), max(X.ncols, 1)))
    return out


def gemm_nn(X: DeviceMatrix, W, alpha=1.0, Y: DeviceMatrix | None = None, beta=0.0) -> DeviceMatrix:
    """Y = alpha * X * W + beta * Y with a small host matrix W."""
    be = backend()
    W = np.asfortranarray(np.asarray(W, dtype=np.float64))
    assert W.shape[0] == X.ncols
    if Y is None:
        Y = DeviceMatrix.empty(W.shape[1])
        beta = 0.0
    assert W.shape[1] == Y.ncols
    be.check(be.lib.dre_gemm_nn(be.h, float(alpha), X.view, capi._dptr(W), max(W.shape[0], 1), float(beta), Y.view))
    _touch(Y)
    return Y


# =============================================================================================
# observers (src/Callbacks.jl:97-187)
# =============================================================================================
# Host-side timer tree with the reference's @timeit_debug section names (TimerOutputs); off unless
# enable_debug_timings(True) -- the analogue of TimerOutputs.enable_debug_timings(DifferentialRiccatiEquations).  The
# library marks the same sections as NVTX ranges (csrc/context.cu: Range).
TIMERS: dict = {}
COUNTS: dict = {}
_DEBUG_TIMINGS = False


def enable_debug_timings(on: bool = True):
    global _DEBUG_TIMINGS
    _DEBUG_TIMINGS = bool(on)
    if on:
        TIMERS.clear()
        COUNTS.clear()


class _timeit:
    __slots__ = ("label", "t0")

    def __init__(self, label):
        self.label = label

    def __enter__(self):
        if _DEBUG_TIMINGS:
            import time

            self.t0 = time.perf_counter()

    def __exit__(self, *exc):
        if _DEBUG_TIMINGS:
            import time

            TIMERS[self.label] = TIMERS.get(self.label, 0.0) + time.perf_counter() - self.t0
            COUNTS[self.label] = COUNTS.get(self.label, 0) + 1
        return False


def _timed(label):
    def deco(fn):
        import functools

        @functools.wraps(fn)
        def wrapper(*a, **kw):
            if not _DEBUG_TIMINGS:
                return fn(*a, **kw)
            with _timeit(label):
                return fn(*a, **kw)
        return wrapper
    return deco


def _observe(observer, name, *args):
    if observer is None:
        return
    fn = getattr(observer, name, None)
    if fn is not None:
        with _timeit("callbacks"):
            fn(*args)


# =============================================================================================
# LDLt (src/LDLt.jl)
# =============================================================================================
class LDLt:
    """src/LDLt.jl:29-33 -- lazy sum_i alpha_i L_i D_i L_i' with device outer factors."""

    def __init__(self, alphas, Ls, Ds):
        self.alphas, self.Ls, self.Ds = list(alphas), list(Ls), list(Ds)

    def destructure(self):  # :54-60
        _join_pending(self)
        if len(self.Ls) > 1:
            compress_(self)
        return self.alphas[0], self.Ls[0], self.Ds[0]

    def __iter__(self):
        return iter(self.destructure())

    @property
    def n(self):
        return self.Ls[0].shape[0]

    @property
    def shape(self):
        return (self.n, self.n)

    def rank(self):  # :112
        _join_pending(self)
        return sum(_ncols(L) for L in self.Ls)

    def to_device_(self):
        """Move host (NumPy) outer factors to the GPU in place; needs an uploaded pencil."""
        for i, L in enumerate(self.Ls):
            if not isinstance(L, DeviceMatrix):
                self.Ls[i] = DeviceMatrix.from_host(L)
        return self

    def iszero(self):  # :114
        return all(a == 0 for a in self.alphas) or sum(_ncols(L) for L in self.Ls) == 0

    def zero(self):  # :116-121
        return LDLt([1.0], [DeviceMatrix.empty(0)], [np.zeros((0, 0))])

    def to_dense(self):  # :42-51 (testing only)
        _join_pending(self)
        M = np.zeros((self.n, self.n))
        for a, L, D in zip(self.alphas, self.Ls, self.Ds):
            Lh = L.to_host()
            M += Lh @ (a * D) @ Lh.T
        return M

    def __add__(self, other):  # :131-148
        if self.iszero():
            return other
        if other.iszero():
            return self
        _join_pending(other)
        out = LDLt(self.alphas + other.alphas, self.Ls + other.Ls, self.Ds + other.Ds)
        p = getattr(self, "_pending", None)
        if p is not None:       # the terms in flight keep their positions at the front of the sum
            out._pending, self._pending = p, None
        return out

    def __neg__(self):  # :150-153
        _join_pending(self)
        return LDLt([-a for a in self.alphas], self.Ls, self.Ds)

    def __sub__(self, other):
        return self + (-other)

    def __rmul__(self, alpha):  # :156-159
        _join_pending(self)
        return LDLt([alpha * a for a in self.alphas], self.Ls, self.Ds)

    def __truediv__(self, alpha):
        return (1.0 / alpha) * self


def lowrank(L, D=None) -> LDLt:
    """src/LDLt.jl:24-27; ``D=None`` is the UniformScaling I.  ``L`` may be a NumPy matrix (moved to the
    GPU when the problem it belongs to is solved) or a DeviceMatrix."""
    Ld = L if isinstance(L, DeviceMatrix) else np.asfortranarray(np.asarray(L, dtype=np.float64))
    Dh = np.eye(_ncols(Ld)) if D is None else np.array(D, dtype=np.float64, order="F")
    return LDLt([1.0], [Ld], [Dh])


def _dcat(Xs, alphas=None):
    """src/util/_dcat.jl:8-22."""
    Xs = list(Xs)
    alphas = [1.0] * len(Xs) if alphas is None else list(alphas)
    n = sum(X.shape[0] for X in Xs)
    D = np.zeros((n, n), order="F")
    k = 0
    for X, a in zip(Xs, alphas):
        l = X.shape[0]
        D[k:k + l, k:k + l] = X * a
        k += l
    return D


@_timed("concatenate!(::LDLt)")
def concatenate_(X: LDLt) -> LDLt:
    """src/LDLt.jl:174-191."""
    _join_pending(X)
    if len(X.alphas) == 1:
        return X
    L = hcat(X.Ls)
    D = _dcat(X.Ds, X.alphas)
    X.alphas[:] = [1.0]
    X.Ls[:] = [L]
    X.Ds[:] = [D]
    return X


def _compress_call(be: Backend, terms):
    """dre_ldlt_compress on context `be`; terms = [(alpha, DeviceMatrix of be, D F-ordered)].  Returns (L, lam)."""
    ktot = sum(L.ncols for _, L, _ in terms)
    nt = len(terms)
    views = (View * nt)(*[L.view for _, L, _ in terms])
    dptrs = (C.POINTER(C.c_double) * nt)(*[capi._dptr(D) for _, _, D in terms])
    ldds = (C.c_int64 * nt)(*[max(D.shape[0], 1) for _, _, D in terms])
    alphas = (C.c_double * nt)(*[float(a) for a, _, _ in terms])
    if _is_orthonormal(terms[0][1]) and np.count_nonzero(terms[0][2] - np.diag(np.diag(terms[0][2]))) == 0:
        be.check(be.lib.dre_hint_orthonormal(be.h, terms[0][1].view))  # outer factor of the previous compress!
    cap = min(ktot, be.n)
    out = be.scratch(cap)  # persistent, geometrically grown: allocating ~2 GB per call costs tens of ms
    lam = np.zeros(cap)
    newrank = C.c_int32(0)
    with _timeit("compress!: dre_ldlt_compress"):
        be.check(be.lib.dre_ldlt_compress(be.h, nt, views, dptrs, ldds, alphas, 100.0, out.view, capi._dptr(lam),
                                          C.byref(newrank)))
    k2 = newrank.value
    with _timeit("compress!: copy out"):
        Lnew = out.cols(0, k2).copy()  # exact-size panel; the scratch panel is reused by the next compress!
    _mark_orthonormal(Lnew)         # Q * (orthonormal eigenvectors): lets the next compress! skip these columns
    return Lnew, lam[:k2]


class CompressStream:
    """compress! in three phases on context `be` (dre_compress_begin / _add / _finish): terms are orthogonalised as
    they are added, finish() runs the core / eigen / L <- QV tail.  Same result as _compress_call on the same terms
    in the same order (identical code path: dre_ldlt_compress is begin + add + finish)."""

    def __init__(self, be: Backend, max_cols: int):
        self.be, self.max_cols, self.cols, self.nterms = be, int(max_cols), 0, 0
        be.check(be.lib.dre_compress_begin(be.h, self.max_cols, 100.0))

    def room_for(self, cols: int) -> bool:
        return self.cols + cols <= self.max_cols

    def scale_hint(self, scale: float):
        """The largest scaled column norm the job will meet (dre_compress_scale_hint); before the first add."""
        self.be.check(self.be.lib.dre_compress_scale_hint(self.be.h, float(scale)))

    def add(self, terms):
        """terms = [(alpha, DeviceMatrix of be, D)]"""
        be = self.be
        terms = [(a, L, np.asfortranarray(D, dtype=np.float64)) for a, L, D in terms if L.ncols]
        if not terms:
            return
        nt = len(terms)
        views = (View * nt)(*[L.view for _, L, _ in terms])
        dptrs = (C.POINTER(C.c_double) * nt)(*[capi._dptr(D) for _, _, D in terms])
        ldds = (C.c_int64 * nt)(*[max(D.shape[0], 1) for _, _, D in terms])
        alphas = (C.c_double * nt)(*[float(a) for a, _, _ in terms])
        if (self.nterms == 0 and _is_orthonormal(terms[0][1])
                and np.count_nonzero(terms[0][2] - np.diag(np.diag(terms[0][2]))) == 0):
            be.check(be.lib.dre_hint_orthonormal(be.h, terms[0][1].view))
        be.check(be.lib.dre_compress_add(be.h, nt, views, dptrs, ldds, alphas))
        self.cols += sum(L.ncols for _, L, _ in terms)
        self.nterms += nt

    def finish(self):
        be = self.be
        cap = max(min(self.cols, be.n), 1)
        out = be.scratch(cap)
        lam = np.zeros(cap)
        newrank = C.c_int32(0)
        be.check(be.lib.dre_compress_finish(be.h, out.view, capi._dptr(lam), C.byref(newrank)))
        k2 = newrank.value
        Lnew = out.cols(0, k2).copy()
        _mark_orthonormal(Lnew)
        return Lnew, lam[:k2]


@_timed("compress!(::LDLt)")
def compress_(X: LDLt) -> LDLt:
    """src/LDLt.jl:204-225 -- one C-ABI call (dre_ldlt_compress); no concatenation copy is needed."""
    _join_pending(X)
    terms = [(a, L, np.asfortranarray(D, dtype=np.float64)) for a, L, D in zip(X.alphas, X.Ls, X.Ds) if L.ncols]
    if sum(L.ncols for _, L, _ in terms) == 0:
        raise ValueError("compress!: rank-0 input (reference: maximum of empty collection, src/LDLt.jl:216)")
    Lnew, lam = _compress_call(backend(), terms)
    _dist.assert_same_int(Lnew.ncols, "the rank after compress!")
    with _timeit("compress!: release terms"):
        del terms
        X.alphas[:] = [1.0]
        X.Ls[:] = [Lnew]
        X.Ds[:] = [np.asfortranarray(np.diag(lam))]
    return X


# Opt-in (DRE_ASYNC_COMPRESS=1): compress!(X) of adi.jl:143-147 on a second context ("lane":
# own stream, workspaces, arena) and a host thread, while the ADI iteration carries on -- the following steps only
# append terms to X (adi.jl:170-176), nothing reads it before the next compression or the end of the solve.
ASYNC_COMPRESS = _os.environ.get("DRE_ASYNC_COMPRESS", "0") not in ("", "0")


LANE_STATS = {"compressions": 0, "lane_busy_s": 0.0, "join_wait_s": 0.0, "start_sync_s": 0.0}   # DRE_ASYNC_COMPRESS


class _PendingCompress:
    """The first `nterms` terms of an LDLt are being compressed on the lane; join() replaces them by the result.
    LANE_STATS accumulates how long the lane worked and how long the ADI thread had to wait for it (wall clock):
    join_wait_s close to lane_busy_s means the two did not overlap."""

    def __init__(self, X: LDLt):
        import threading
        import time

        t0 = time.perf_counter()

        main = backend()
        self.lane = _lane(main)
        main.ctx.sync()                      # every kernel that produced the terms has finished
        LANE_STATS["start_sync_s"] += time.perf_counter() - t0
        LANE_STATS["compressions"] += 1
        keep = [(a, L, np.asfortranarray(D, dtype=np.float64)) for a, L, D in zip(X.alphas, X.Ls, X.Ds)]
        self.nterms = len(keep)
        self.src = keep                      # the main context's panels stay alive until join()
        terms = []
        for a, L, D in keep:
            if L.ncols == 0:
                continue
            W = _alias(self.lane, L)
            if _is_orthonormal(L):
                _mark_orthonormal(W)
            terms.append((a, W, D))
        self.terms, self.result, self.error = terms, None, None
        self.thread = threading.Thread(target=self._run, name="dre-compress-lane", daemon=True)
        self.thread.start()

    def _run(self):
        import time

        t0 = time.perf_counter()
        try:
            res = _compress_call(self.lane, self.terms)
            self.lane.ctx.sync()
            self.result = res
        except BaseException as e:           # re-raised by join() on the caller's thread
            self.error = e
        LANE_STATS["lane_busy_s"] += time.perf_counter() - t0

    def join(self, X: LDLt):
        import time

        t0 = time.perf_counter()
        self.thread.join()
        LANE_STATS["join_wait_s"] += time.perf_counter() - t0
        self.terms = None                    # lane-side aliases of the inputs (the lane is idle now)
        if self.error is not None:
            raise self.error
        Llane, lam = self.result
        Lnew = _alias(backend(), Llane)      # the compressed factor stays in the lane's arena
        _mark_orthonormal(Lnew)
        X.alphas[:self.nterms] = [1.0]
        X.Ls[:self.nterms] = [Lnew]
        X.Ds[:self.nterms] = [np.asfortranarray(np.diag(lam))]
        self.src = None


class _RemotePending:
    """Multi-GPU pipeline mode: the terms of X have been streamed to rank 1, which holds and compresses the sum;
    join() fetches the compressed factor (dre_b200.dist.pipe_fetch)."""

    def join(self, X: LDLt):
        # every term of X was streamed right after the step that produced it, so rank 1's X is the whole sum
        be = backend()
        Lnew, lam = _dist.pipe_fetch(be, lambda k: DeviceMatrix.empty(k))
        _mark_orthonormal(Lnew)
        X.alphas[:] = [1.0]
        X.Ls[:] = [Lnew]
        X.Ds[:] = [np.asfortranarray(np.diag(lam))]


def _join_pending(X: LDLt):
    p = getattr(X, "_pending", None)
    if p is not None:
        X._pending = None
        p.join(X)


@_timed("norm(::LDLt)")
def norm(X: LDLt) -> float:
    """src/LDLt.jl:77-89."""
    be = backend()
    X.to_device_()
    concatenate_(X)
    a, L, D = X.alphas[0], X.Ls[0], X.Ds[0]
    if L.ncols == 0:
        return 0.0
    D = np.asfortranarray(D, dtype=np.float64)
    out = C.c_double(0.0)
    be.check(be.lib.dre_ldlt_norm(be.h, L.view, capi._dptr(D), max(D.shape[0], 1), float(a), C.byref(out)))
    return float(out.value)


@_timed("dot(::LDLt, ::LDLt)")
def dot(X1: LDLt, X2: LDLt) -> float:
    """src/LDLt.jl:91-108 -- Frobenius inner product tr(X1' X2): one Gram product A'C on the device
    (dre_gemm_tn), the k1 x k2 core algebra on the host."""
    if X1.n != X2.n:
        raise ValueError("DimensionMismatch")
    X1.to_device_()
    X2.to_device_()
    concatenate_(X1)
    concatenate_(X2)
    alpha, A, B = X1.alphas[0], X1.Ls[0], np.asarray(X1.Ds[0])
    beta, Cm, D = X2.alphas[0], X2.Ls[0], np.asarray(X2.Ds[0])
    if A.ncols == 0 or Cm.ncols == 0:
        return 0.0
    AtC = gemm_tn(A, Cm)
    M = (B.T @ AtC @ D) * (alpha * beta)
    return float(np.sum(AtC * M))


# =============================================================================================
# operators: pencil combinations and LowRankUpdate (src/LowRankUpdate.jl)
# =============================================================================================
class PencilCombo:
    """The sparse matrix a*A + e*E of the uploaded (symmetric) pencil -- what the reference builds
    as a new SparseMatrixCSC per shift (adi.jl:156, lowrank_ros1.jl:39) is two scalars here."""

    def __init__(self, a: float, e: float):
        self.a, self.e = float(a), float(e)

    @property
    def shape(self):
        n = backend().n
        return (n, n)

    @property
    def T(self):
        return self  # symmetric pencil

    def __add__(self, other):
        if isinstance(other, PencilCombo):
            return PencilCombo(self.a + other.a, self.e + other.e)
        return NotImplemented

    def __sub__(self, other):
        return PencilCombo(self.a - other.a, self.e - other.e)

    def __mul__(self, s):
        return PencilCombo(self.a * s, self.e * s)

    __rmul__ = __mul__

    def __truediv__(self, s):
        return PencilCombo(self.a / s, self.e / s)

    def matmul(self, X: DeviceMatrix) -> DeviceMatrix:
        Y = None
        if self.a != 0.0 or self.e == 0.0:
            Y = spmm("A", X, self.a)
        if self.e != 0.0:
            Y = spmm("E", X, self.e, Y, 1.0 if Y is not None else 0.0)
        return Y


class LowRankUpdate:
    """src/LowRankUpdate.jl:18-26 -- lazy A + inv(alpha) U V with U (n x m) and V (m x n, held as its
    transpose panel Vt, n x m)."""

    def __init__(self, A: PencilCombo, alpha: float, U: DeviceMatrix, Vt: DeviceMatrix):
        self.A, self.alpha, self.U, self.Vt = A, float(alpha), U, Vt

    @property
    def shape(self):
        return self.A.shape

    def adjoint(self):  # :51-54
        return LowRankUpdate(self.A.T, self.alpha, self.Vt, self.U)

    def plus_sparse(self, E: PencilCombo):  # :66-70
        return LowRankUpdate(self.A + E, self.alpha, self.U, self.Vt)

    def matmul(self, X: DeviceMatrix) -> DeviceMatrix:  # :72-86
        if X.ncols == X.shape[0]:
            warnings.warn("Multiplying LowRankUpdate by square matrix; memory usage may increase severely")
        Y = self.A.matmul(X)
        VX = gemm_tn(self.Vt, X)  # m x k
        return gemm_nn(self.U, VX, 1.0 / self.alpha, Y, 1.0)


def lr_update(A, alpha, U, V_or_Vt, *, transposed=False):
    """src/LowRankUpdate.jl:38-39.  ``V`` is m x n on the host, or (transposed=True) its n x m panel."""
    if not isinstance(A, PencilCombo):
        raise TypeError("lr_update: A must be a PencilCombo of the uploaded pencil")
    Ud = _as_device(U)
    Vt = V_or_Vt if transposed else _as_device(np.asarray(V_or_Vt).T)
    return LowRankUpdate(A, alpha, Ud, Vt)


def _matmul(A, X: DeviceMatrix) -> DeviceMatrix:
    return A.matmul(X)


def _adj_matmul(A, X: DeviceMatrix) -> DeviceMatrix:
    return A.adjoint().matmul(X) if isinstance(A, LowRankUpdate) else A.T.matmul(X)


# =============================================================================================
# block linear solvers (src/blocklinear/*.jl)
# =============================================================================================
class BlockLinearProblem:  # types.jl:10-13
    def __init__(self, A, B):
        self.A, self.B = A, B


class BlockLinearSolver:
    pass


class Backslash(BlockLinearSolver):  # types.jl:31-34
    """Sparse direct solve on the GPU: numeric supernodal LDL^T per shift + block sweeps."""


class ShermanMorrisonWoodbury(BlockLinearSolver):  # types.jl:35-39
    def __init__(self, ALG=None, alg=None):
        self.ALG = ALG if ALG is not None else Backslash()
        self.alg = alg if alg is not None else Backslash()


def _set_operator(F, *, transpose: bool):
    """Tell the library which operator the next shifted solves use.  The library solves with
    (F' + mu E'); ``transpose=False`` (solve with F itself) swaps the low-rank factors."""
    be = backend()
    if isinstance(F, LowRankUpdate):
        U, Vt = (F.U, F.Vt) if transpose else (F.Vt, F.U)
        be.check(be.lib.dre_set_operator(be.h, F.A.a, F.A.e, F.alpha, U.view, Vt.view))
    else:
        z = View(-1, 0, 0)
        be.check(be.lib.dre_set_operator(be.h, F.a, F.e, 1.0, z, z))


def solve_block(prob: BlockLinearProblem, alg: BlockLinearSolver | None = None, *, mu: complex = 0.0):
    """CommonSolve.solve(::BlockLinearProblem, alg): X = (prob.A + mu E) \\ prob.B  (backslash.jl:8-21,
    sherman-morrison-woodbury.jl:10-45 when prob.A is a LowRankUpdate).  Returns a DeviceMatrix, or a
    pair (Re X, Im X) for complex mu."""
    be = backend()
    alg = alg if alg is not None else Backslash()
    if not isinstance(alg, (Backslash, ShermanMorrisonWoodbury)):
        return alg.solve(prob, mu=mu)  # user-defined BlockLinearSolver (types.jl:15-30)
    _set_operator(prob.A, transpose=False)
    B = prob.B
    V1 = DeviceMatrix.empty(B.ncols)
    mu = complex(mu)
    if mu.imag == 0.0:
        be.check(be.lib.dre_shift_solve(be.h, mu.real, 0.0, B.view, V1.view, View(-1, 0, 0)))
        return V1
    V2 = DeviceMatrix.empty(B.ncols)
    be.check(be.lib.dre_shift_solve(be.h, mu.real, mu.imag, B.view, V1.view, V2.view))
    return V1, V2


# =============================================================================================
# Shifts (src/Shifts.jl, src/shifts/*.jl)
# =============================================================================================
class Shifts:
    """Namespace mirroring the reference's ``Shifts`` module."""

    class Strategy:
        pass

    @staticmethod
    def safe_sort(shifts):  # helpers.jl:122
        shifts = list(shifts)
        shifts.sort(key=lambda v: (np.real(v), abs(np.imag(v))))
        return shifts

    @staticmethod
    def stabilize_ritz_values(lam, desc):  # helpers.jl:129-140
        lam = list(lam)
        assert len(lam) > 0
        n_unstable = sum(1 for v in lam if not np.real(v) < 0)
        if 0 < n_unstable < len(lam):
            warnings.warn(f"Discarding unstable Ritz values of {desc}")
            lam = [v for v in lam if np.real(v) < 0]
        elif n_unstable == len(lam):
            warnings.warn(f"All Ritz values of {desc} are unstable; flipping along imaginary axis")
            lam = [complex(-np.real(v), np.imag(v)) if np.iscomplexobj(v) else -v for v in lam]
        return lam

    @staticmethod
    def heuristic(R, nshifts=None):  # heuristic.jl:82-101
        R = list(R)
        nshifts = len(R) if nshifts is None else nshifts

        def s(t, P):
            return math.prod(abs(t - p) / abs(t + p) for p in P)

        vals = [max(s(t, (p,)) for t in R) for p in R]
        p = R[int(np.argmin(vals))]
        P = [p] if np.imag(p) == 0 else [p, np.conj(p)]
        while len(P) < nshifts:
            vals = [s(t, P) for t in R]
            p = R[int(np.argmax(vals))]
            if np.imag(p) == 0:
                P.append(p)
            else:
                P.extend((p, np.conj(p)))
        return P


class Projection(Shifts.Strategy):  # projection.jl:25-32
    def __init__(self, u):
        if u % 2 == 1:
            raise ValueError(f"History must be even; got {u}")
        self.n_history = u


class Heuristic(Shifts.Strategy):  # heuristic.jl:22-30
    def __init__(self, nshifts, k_plus, k_minus, alg_E=None, alg_A=None):
        self.nshifts, self.k_plus, self.k_minus = nshifts, k_plus, k_minus
        self.alg_E = alg_E if alg_E is not None else Backslash()
        self.alg_A = alg_A if alg_A is not None else Backslash()


class Cyclic(Shifts.Strategy):  # helpers.jl:19-27
    def __init__(self, inner):
        self.inner = inner


class Wrapped(Shifts.Strategy):  # helpers.jl:48-58
    def __init__(self, func, inner):
        self.func, self.inner = func, inner


Shifts.Projection, Shifts.Heuristic, Shifts.Cyclic, Shifts.Wrapped = Projection, Heuristic, Cyclic, Wrapped


ORTH_TRACE = None  # set to a list to record (singular values, kept directions) of every orth_restrict call


def orth_restrict(Vs, E, A):
    """Q = orth(hcat(Vs)) (src/Stuff.jl:13-18) and the restrictions Q'EQ, Q'AQ (src/Stuff.jl:9,
    src/util/restrict.jl:5-8), without ever forming the SVD basis on the device: a rank-revealing QR
    N = Q0 Rt' runs on the GPU, the small SVD of Rt' on the host selects the singular directions
    > n*eps exactly as the reference, and the projected pencils are rotated on the host."""
    be = backend()
    n = be.n
    ktot = sum(V.ncols for V in Vs)
    if ktot == 0:
        return np.zeros((0, 0)), np.zeros((0, 0))
    views = (View * len(Vs))(*[V.view for V in Vs])
    cap = min(ktot, n)
    Q0 = DeviceMatrix.empty(cap)
    Rt = np.zeros((ktot, cap), order="F")
    rho = C.c_int32(0)
    be.check(be.lib.dre_rrqr(be.h, len(Vs), views, 1e-15, 1e-3 * n * EPS, Q0.view, capi._dptr(Rt), ktot,
                             C.byref(rho)))
    rho = rho.value
    if rho == 0:
        return np.zeros((0, 0)), np.zeros((0, 0))
    Q0 = Q0.cols(0, rho)
    U, s, _ = sla.svd(Rt[:, :rho].T, full_matrices=False, lapack_driver="gesdd")  # N = Q0 (U s W')
    ids = np.nonzero(np.abs(s) > n * EPS)[0]  # Stuff.jl:15-16
    if ORTH_TRACE is not None:  # diagnostics of the free-run tests: singular values and kept count of this refill
        ORTH_TRACE.append(dict(k=ktot, rho=rho, kept=len(ids), s=s.copy()))
    Us = U[:, ids]
    EQ = spmm("E", Q0)
    Et = gemm_tn(Q0, EQ)
    if isinstance(A, LowRankUpdate):
        AQ = A.A.matmul(Q0)
        At = gemm_tn(Q0, AQ) + (1.0 / A.alpha) * (gemm_tn(Q0, A.U) @ gemm_tn(A.Vt, Q0))
    else:
        At = gemm_tn(Q0, A.matmul(Q0))
    return Us.T @ Et @ Us, Us.T @ At @ Us


def _pencil_eigvals(At, Et):
    """eigvals(At, Et) of src/shifts/projection.jl:67.  The projected mass matrix Q'EQ is symmetric positive
    definite for the FEM pencils of this path, so the pencil is reduced to the standard problem
    C^-1 At C^-T (Et = C C') and solved with the QR algorithm (LAPACK geev) -- about 3x cheaper on the host than
    the QZ algorithm behind eigvals(A, B) for the ~500 x 500 pencils met here; same Ritz values up to round-off.
    Falls back to the generalized solver when Et is not numerically SPD."""
    if Et.shape[0] == 0:
        return np.zeros(0)
    try:
        if not np.allclose(Et, Et.T, rtol=1e-10, atol=1e-14 * np.abs(Et).max()):
            raise np.linalg.LinAlgError
        Cf = np.linalg.cholesky(0.5 * (Et + Et.T))
        M = sla.solve_triangular(Cf, At, lower=True)
        M = sla.solve_triangular(Cf, M.T, lower=True).T
        return sla.eigvals(M)
    except np.linalg.LinAlgError:
        return sla.eigvals(At, Et)


class ProjectionShiftIterator:  # projection.jl:34-73
    def __init__(self, prob, n_history):
        self.prob, self.n_history, self.Vs = prob, n_history, []

    def update(self, X, R, *Vs):  # :45-52
        if not Vs:
            self.Vs.append(R)  # aliases the residual factor, which ADI updates in place (adi.jl:171)
        self.Vs.extend(Vs)
        lst = len(self.Vs)
        fst = max(0, lst - self.n_history)
        self.Vs = self.Vs[fst:lst]

    def take_many(self):  # :54-73
        Et, At = orth_restrict(self.Vs, self.prob.E, self.prob.A)
        lam = _pencil_eigvals(At, Et)
        if np.all(np.imag(lam) == 0):
            lam = np.real(lam)
        lam = Shifts.stabilize_ritz_values(lam, "(A, E)")
        return Shifts.safe_sort(lam)


class BufferedIterator:  # helpers.jl:70-75, 106-113
    def __init__(self, gen):
        self.buffer, self.generator = [], gen

    def update(self, *args):
        upd = getattr(self.generator, "update", None)
        if upd is not None:
            upd(*args)

    def take(self):
        if not self.buffer:
            self.buffer = list(_dist.agree_array(np.asarray(self.generator.take_many())))
        return self.buffer.pop(0)

    def peek_many(self, k):
        """Up to k next shifts that are already buffered (never triggers take_many!)."""
        return list(self.buffer[:k])


class WrappedIterator:  # helpers.jl:85-104
    def __init__(self, func, gen):
        self.func, self.generator = func, gen

    def update(self, *args):
        upd = getattr(self.generator, "update", None)
        if upd is not None:
            upd(*args)

    def take_many(self):
        return self.func(_take_many(self.generator))


class _CycleIterator:  # Stateful(cycle(values)), helpers.jl:93
    def __init__(self, values):
        self.values = list(values)
        self._it = itertools.cycle(self.values)

    def update(self, *args):
        return None

    def take(self):
        return next(self._it)

    def peek_many(self, k):
        self._it, probe = itertools.tee(self._it)
        return list(itertools.islice(probe, k))


class _ListIterator:  # plain vector: take! = popfirst! (Shifts.jl:116)
    def __init__(self, values):
        self.values = list(values)

    def update(self, *args):
        return None

    def take(self):
        return self.values.pop(0)

    def peek_many(self, k):
        return list(self.values[:k])

    def take_many(self):
        return self.values


def _take_many(gen):
    return gen.take_many() if hasattr(gen, "take_many") else list(gen)


def _arnoldi_ritz(op, b0, k, desc):
    """heuristic.jl:103-130 -- Arnoldi with twice-repeated MGS.  The Krylov basis lives in ONE device panel
    (n x (k+1)); ``op`` maps a device column to a device column (SpMM + shifted solve); every orthogonalisation
    step is one ``dre_arnoldi_orth`` call (chain of dot/update launches, k+2 doubles back).  Only H is on the host,
    like the reference keeps it (``H = zeros(k + 1, k)``)."""
    be = backend()
    H = np.zeros((k + 1, k))
    V = DeviceMatrix.empty(k + 1)
    v0 = np.asfortranarray(((1.0 / np.linalg.norm(b0)) * np.asarray(b0, dtype=np.float64)).reshape(-1, 1))
    be.check(be.lib.dre_mat_upload(be.h, V.cols(0, 1).view, capi._dptr(v0), v0.shape[0]))
    h = np.zeros(k + 2)
    for j in range(k):
        w = op(V.cols(j, j + 1))
        be.check(be.lib.dre_arnoldi_orth(be.h, V.cols(0, j + 1).view, w.view, V.cols(j + 1, j + 2).view,
                                         capi._dptr(h)))
        H[:j + 2, j] = h[:j + 2]
    ritz = sla.eigvals(H[:k, :k])
    if np.all(np.imag(ritz) == 0):
        ritz = np.real(ritz)
    return Shifts.stabilize_ritz_values(ritz, desc)


def shifts_init(strategy, prob):
    """Shifts.init (projection.jl:40-43, heuristic.jl:39-66, helpers.jl:86-99)."""
    if isinstance(strategy, Projection):
        return BufferedIterator(ProjectionShiftIterator(prob, strategy.n_history))
    if isinstance(strategy, Heuristic):
        E, A = prob.E, prob.A
        n = backend().n
        b0 = np.ones(n)  # heuristic.jl:75-80

        def op_plus(x):  # E \ (A x), device column -> device column
            return solve_block(BlockLinearProblem(E, _matmul(A, x)), strategy.alg_E)

        def op_minus(x):  # A \ (E x)
            return solve_block(BlockLinearProblem(A, _matmul(E, x)), strategy.alg_A)

        R_plus = _arnoldi_ritz(op_plus, b0, strategy.k_plus, "E^-1 A")
        R_minus = _arnoldi_ritz(op_minus, b0, strategy.k_minus, "A^-1 E")
        R = list(R_plus) + [1.0 / v for v in R_minus]
        return _ListIterator(Shifts.heuristic(R, strategy.nshifts))
    if isinstance(strategy, Cyclic):
        inner = strategy.inner
        vals = _take_many(shifts_init(inner, prob)) if isinstance(inner, Shifts.Strategy) else list(inner)
        return _CycleIterator(vals)
    if isinstance(strategy, Wrapped):
        it = shifts_init(strategy.inner, prob)
        if isinstance(it, BufferedIterator):
            return BufferedIterator(WrappedIterator(strategy.func, it.generator))
        return WrappedIterator(strategy.func, it)
    return strategy.init(prob)  # custom strategy protocol (Shifts.jl:13-67)


# =============================================================================================
# GALE / ADI (src/lyapunov/types.jl, residual.jl, adi.jl)
# =============================================================================================
class GALEProblem:  # types.jl:10-16   A'XE + E'XA = -C
    def __init__(self, E, A, C):
        self.E, self.A, self.C = E, A, C


class ADI:  # types.jl:20-32
    def __init__(self, inner_alg=None, *, maxiters=100, reltol=None, abstol=None, shifts=None,
                 ignore_initial_guess=False, compression_interval=10, compression=True, warn_convergence=True):
        self.maxiters, self.reltol, self.abstol = maxiters, reltol, abstol
        self.shifts = shifts if shifts is not None else Projection(2)
        self.ignore_initial_guess = ignore_initial_guess
        self.inner_alg = inner_alg if inner_alg is not None else Backslash()
        self.compression_interval, self.compression = compression_interval, compression
        self.warn_convergence = warn_convergence


def _device_problem(prob: GALEProblem) -> GALEProblem:
    """Accept host inputs (scipy E/A, NumPy factors) like the reference accepts SparseMatrixCSC/Matrix
    and move them to the GPU once."""
    E, A, Cm = prob.E, prob.A, prob.C
    if sp.issparse(E):
        if not sp.issparse(A):
            raise TypeError("E is sparse but A is not")
        backend().ensure_pencil(E, A)
        E, A = PencilCombo(0.0, 1.0), PencilCombo(1.0, 0.0)
    if not isinstance(E, PencilCombo) or not isinstance(A, (PencilCombo, LowRankUpdate)):
        raise TypeError("GALEProblem: E/A must be scipy sparse matrices or device operators")
    if (E.a, E.e) != (0.0, 1.0):
        raise ValueError("GALEProblem: E must be the uploaded mass matrix")
    Cm.to_device_()
    return GALEProblem(E, A, Cm)


@_timed("residual(::GALEProblem, ::LDLt)")
def residual(prob: GALEProblem, val: LDLt) -> LDLt:
    """src/lyapunov/residual.jl:3-31."""
    E, A, Cm = prob.E, prob.A, prob.C
    if val.iszero():
        return LDLt(list(Cm.alphas), [L.copy() for L in Cm.Ls], [np.array(D, order="F") for D in Cm.Ds])
    alpha, G, S = Cm.destructure()
    beta, L, D = val.destructure()
    n_G, n_0 = G.ncols, L.ncols
    dim = n_G + 2 * n_0
    be = backend()
    R = DeviceMatrix.empty(dim)
    be.check(be.lib.dre_mat_copy(be.h, R.cols(0, n_G).view, G.view))
    spmm("E", L, 1.0, R.cols(n_G, n_G + n_0), 0.0)
    AtL = _adj_matmul(A, L)
    be.check(be.lib.dre_mat_copy(be.h, R.cols(n_G + n_0, dim).view, AtL.view))
    T = np.zeros((dim, dim), order="F")
    T[:n_G, :n_G] = alpha * S
    T[n_G:n_G + n_0, n_G + n_0:] = beta * D
    T[n_G + n_0:, n_G:n_G + n_0] = T[n_G:n_G + n_0, n_G + n_0:]
    return compress_(LDLt([1.0], [R], [T]))


class ADICache:
    """src/lyapunov/adi.jl:5-21."""

    def __init__(self, **kw):
        self.last_compression = 0
        self.__dict__.update(kw)

    def __iter__(self):  # adi.jl:91-95
        done = False
        while not done:
            step_(self)
            done = isdone(self)
            yield self


def init(prob: GALEProblem, alg: ADI, *, initial_guess=None, initial_residual=None, abstol=None,
         observer=None) -> ADICache:
    """CommonSolve.init(::GALEProblem{<:LDLt}, ::ADI) -- src/lyapunov/adi.jl:29-69."""
    _observe(observer, "observe_gale_start", prob, alg)
    prob = _device_problem(prob)
    Cm = prob.C
    if alg.ignore_initial_guess or initial_guess is None:
        initial_guess = Cm.zero()
    initial_guess.to_device_()
    if initial_residual is None:
        initial_residual = residual(prob, initial_guess)
    X = initial_guess
    _, R, _T = initial_residual.destructure()
    residual_norm = norm(initial_residual)
    oracle = shifts_init(alg.shifts, prob)
    oracle.update(X, R)
    reltol = alg.reltol if alg.reltol is not None else prob.A.shape[0] * EPS
    if abstol is None:
        abstol = alg.abstol if alg.abstol is not None else reltol * norm(Cm)
    _observe(observer, "observe_gale_step", 0, X, initial_residual, residual_norm)
    increment = initial_residual.zero()
    cache = ADICache(prob=prob, alg=alg, abstol=abstol, observer=observer, shifts_oracle=oracle, shifts=[], X=X,
                     increment=increment, residual=initial_residual, residual_norm=residual_norm)
    cache._piped = bool(_dist.pipeline_active() and alg.compression and _fused_inner(alg))
    if cache._piped:
        _join_pending(X)
        _dist.pipe_begin(backend(), X)   # rank 1 holds the terms of X from now on (marked on cache.X at the first step)
    return cache


def isdone(cache: ADICache) -> bool:
    """adi.jl:130-141."""
    if cache.residual_norm <= cache.abstol:
        return True
    niters = len(cache.shifts)
    if niters > 0 and cache.increment.iszero():
        return True
    return niters >= cache.alg.maxiters


def compress_cache_(cache: ADICache):
    """adi.jl:143-147."""
    if getattr(cache, "_piped", False):
        # pipeline mode: every term of X has already been streamed to rank 1; it compresses there while this
        # rank's iteration carries on (nothing reads X before the next compression or the end of the solve)
        _dist.pipe_compress()
        cache.last_compression = 0
        return
    compress_(cache.X)
    cache.last_compression = 0


def _start_async_compress(cache: ADICache):
    """compress_cache_ on the compression lane (DRE_ASYNC_COMPRESS): at most one compression is in flight."""
    X = cache.X
    _join_pending(X)
    X._pending = _PendingCompress(X)
    cache.last_compression = 0


def _pipe_stream_increment(cache: ADICache):
    """Pipeline mode (dre_b200.dist): the new terms of X go to rank 1 as soon as they exist (asynchronous NCCL send
    queued behind the solve on the library's stream)."""
    if not getattr(cache, "_piped", False):
        return
    if not isinstance(getattr(cache.X, "_pending", None), _RemotePending):
        cache.X._pending = _RemotePending()      # (X + increment returned a fresh object: X was zero)
    be = backend()
    inc = cache.increment
    for a, L, D in zip(inc.alphas, inc.Ls, inc.Ds):
        if L.ncols:
            _dist.pipe_send_term(be, a, L, D)


def _fused_inner(alg: ADI) -> bool:
    return isinstance(alg.inner_alg, (Backslash, ShermanMorrisonWoodbury))


# factorizations queued ahead of the ADI step in flight (the library keeps 4 factor slots); DRE_PREFACTOR_DEPTH=0..3
PREFACTOR_DEPTH = max(0, min(3, int(_os.environ.get("DRE_PREFACTOR_DEPTH", "3"))))


def _prefetch_next_factorization(cache: ADICache):
    """Performance hint only: the next shifts are usually known (buffered), so their numeric factorizations
    are queued on the library's side streams and overlap this step's remaining work."""
    peek = getattr(cache.shifts_oracle, "peek_many", None)
    if peek is None or not _fused_inner(cache.alg):
        return
    left = cache.alg.maxiters - len(cache.shifts)
    be = backend()
    queued, skip = 0, False
    for nxt in peek(2 * PREFACTOR_DEPTH):
        if left <= 0 or queued >= PREFACTOR_DEPTH:
            break
        left -= 1
        if skip:  # conjugate partner of a complex shift: the double step needs one factorization only
            skip = False
            continue
        nxt = complex(nxt)
        skip = nxt.imag != 0
        be.check(be.lib.dre_prefactor(be.h, nxt.real, nxt.imag))
        queued += 1


# Overlap of the residual norm with the next shifted solve.  The
# reference computes norm(residual) after every ADI step only to test for convergence (adi.jl:118-127); the Gram
# product behind it keeps the whole GPU busy for ~0.4 ms while the sweeps of the next step are latency bound.  With
# the option on, step_ queues the norm on a side stream (dre_ldlt_norm_begin), starts the solve half of the NEXT step
# from the already buffered shift while it runs (R is only read), and then collects the norm.  If ADI stops, the
# speculative block is dropped; otherwise the next step adopts it and only runs its residual update.  Results are
# identical: the same kernels run on the same data.
# Default since run r02v (1.331 -> 1.356 steps/s at n = 79 841); DRE_ASYNC_NORM=0 restores the synchronous norm.  Single-GPU
# runs only: the multi-GPU modes keep the plain sequence (their exchange steps are ordered behind the norm).
ASYNC_NORM = _os.environ.get("DRE_ASYNC_NORM", "1") not in ("", "0")


def _speculate_next_solve(cache: ADICache):
    """Queue V = (F' + mu E')^-1 R for the next buffered shift; returns (mu, V1, V2 or None) or None."""
    peek = getattr(cache.shifts_oracle, "peek_many", None)
    if peek is None or not _fused_inner(cache.alg) or _dist.active():
        return None
    if len(cache.shifts) >= cache.alg.maxiters:
        return None
    nxt = peek(2)
    if not nxt:
        return None
    mu = complex(nxt[0])
    if mu.imag != 0 and (len(nxt) < 2 or not np.isclose(complex(nxt[1]), np.conj(mu))):
        return None
    be = backend()
    _, R, _T = cache.residual.destructure()
    _set_operator(cache.prob.A, transpose=True)
    V1 = DeviceMatrix.empty(R.ncols)
    V2 = DeviceMatrix.empty(R.ncols) if mu.imag != 0 else None
    be.check(be.lib.dre_adi_solve(be.h, mu.real, mu.imag, R.view, V1.view,
                                  V2.view if V2 is not None else View(-1, 0, 0)))
    return (mu, V1, V2)


def _adopt_speculation(cache: ADICache, mu: complex):
    """The block solved ahead of time for this shift, if any (and if the residual factor is still the same panel)."""
    spec = getattr(cache, "_spec", None)
    cache._spec = None
    if spec is None:
        return None
    smu, V1, V2, Rid = spec
    _, R, _T = cache.residual.destructure()
    if smu != complex(mu) or Rid != (R.view.id, R.view.col0, R.view.ncols):
        return None
    cache.adopted_speculations = getattr(cache, "adopted_speculations", 0) + 1
    return V1, V2


def perform_single_step_(cache: ADICache, mu: float):
    """adi.jl:149-179: V = (A' + mu E')^-1 R; X += -2 mu alpha V T V'; R -= 2 mu E' V."""
    be = backend()
    prob, alg = cache.prob, cache.alg
    alpha, R, T = cache.residual.destructure()
    adopted = _adopt_speculation(cache, complex(mu)) if _fused_inner(alg) else None
    if adopted is not None:
        V = adopted[0]
        spmm("E", V, -2.0 * mu, R, 1.0)
        _prefetch_next_factorization(cache)
    elif _fused_inner(alg):
        # one C-ABI call: numeric LDL^T of A_s + mu E, block sweeps for [R, K'], SMW, SpMM update
        _set_operator(prob.A, transpose=True)
        V = DeviceMatrix.empty(R.ncols)
        if _dist.active():
            # column blocks of R on different GPUs, replicated factorization; all-gather of V, then the
            # residual update on the full panel (cheaper than gathering the updated R blocks as well)
            _dist.sharded_adi_solve(be, complex(mu), R, V, None, View(-1, 0, 0))
            spmm("E", V, -2.0 * mu, R, 1.0)
        else:
            be.check(be.lib.dre_adi_step(be.h, mu, 0.0, R.view, V.view, View(-1, 0, 0)))
            _touch(R)
        _prefetch_next_factorization(cache)
    else:
        F = (prob.A.adjoint().plus_sparse(PencilCombo(0.0, mu)) if isinstance(prob.A, LowRankUpdate)
             else prob.A.T + PencilCombo(0.0, mu))
        V = alg.inner_alg.solve(BlockLinearProblem(F, R))
        spmm("E", V, -2.0 * mu, R, 1.0)
    cache.increment = (-2.0 * mu * alpha) * LDLt([1.0], [V], [T])
    cache.X = cache.X + cache.increment
    _pipe_stream_increment(cache)
    cache.last_compression += 1
    cache.shifts_oracle.update(cache.X, R, V)


def perform_double_step_(cache: ADICache, mu: complex):
    """adi.jl:181-225 (complex pair; note (conj(mu) E)' == mu E', adi.jl:195)."""
    be = backend()
    prob, alg = cache.prob, cache.alg
    alpha, R, T = cache.residual.destructure()
    mu_next = cache.shifts_oracle.take()
    assert np.isclose(mu_next, np.conj(mu)), (mu, mu_next)
    cache.shifts.append(complex(mu_next))
    _observe(cache.observer, "observe_gale_metadata", "ADI shifts", mu_next)
    adopted = _adopt_speculation(cache, complex(mu)) if _fused_inner(alg) else None
    if adopted is not None:
        V1, V2 = adopted
        spmm("E", V1, -2.0 * math.sqrt(2.0) * mu.real, R, 1.0)
        _prefetch_next_factorization(cache)
    elif _fused_inner(alg):
        _set_operator(prob.A, transpose=True)
        V1 = DeviceMatrix.empty(R.ncols)
        V2 = DeviceMatrix.empty(R.ncols)
        if _dist.active():
            _dist.sharded_adi_solve(be, complex(mu), R, V1, V2, View(-1, 0, 0))
            spmm("E", V1, -2.0 * math.sqrt(2.0) * mu.real, R, 1.0)
        else:
            be.check(be.lib.dre_adi_step(be.h, mu.real, mu.imag, R.view, V1.view, V2.view))
            _touch(R)
        _prefetch_next_factorization(cache)
    else:
        F = (prob.A.adjoint().plus_sparse(PencilCombo(0.0, 0.0)) if isinstance(prob.A, LowRankUpdate)
             else prob.A.T)
        Vr, Vi = alg.inner_alg.solve(BlockLinearProblem(F, R), mu=mu)
        d = mu.real / mu.imag
        V1 = Vr.copy()
        be.check(be.lib.dre_mat_axpby(be.h, math.sqrt(2.0) * d, Vi.view, math.sqrt(2.0), V1.view))
        _touch(V1)
        V2 = Vi.copy()
        be.check(be.lib.dre_mat_axpby(be.h, 0.0, View(-1, 0, 0), math.sqrt(2 * d * d + 2), V2.view))
        _touch(V2)
        spmm("E", V1, -2.0 * math.sqrt(2.0) * mu.real, R, 1.0)
    cache.increment = (-2.0 * mu.real * alpha) * (LDLt([1.0], [V1], [T]) + LDLt([1.0], [V2], [T]))
    cache.X = cache.X + cache.increment
    _pipe_stream_increment(cache)
    cache.last_compression += 2
    cache.shifts_oracle.update(cache.X, R, V1, V2)


def _residual_norm_overlapped(cache: ADICache) -> float:
    """norm(cache.residual); with DRE_ASYNC_NORM=1 the next step's solve is queued while the norm is computed."""
    X = cache.residual
    if (not ASYNC_NORM or len(X.alphas) != 1 or X.Ls[0].ncols == 0 or _dist.active()
            or getattr(cache, "_piped", False)):
        return norm(X)
    D = np.asarray(X.Ds[0], dtype=np.float64)
    if np.count_nonzero(D - np.diag(np.diag(D))) != 0:
        return norm(X)          # dense core: the synchronous path finishes on the host
    be = backend()
    a, L = X.alphas[0], X.Ls[0]
    d = np.ascontiguousarray(np.diag(D))
    be.check(be.lib.dre_ldlt_norm_begin(be.h, L.view, capi._dptr(d), float(a)))
    try:
        spec = _speculate_next_solve(cache)
        cache._spec = None if spec is None else (spec[0], spec[1], spec[2], (L.view.id, L.view.col0, L.view.ncols))
    finally:
        out = C.c_double(0.0)
        be.check(be.lib.dre_ldlt_norm_end(be.h, C.byref(out)))
    return float(out.value)


def step_(cache: ADICache):
    """CommonSolve.step!(::ADICache) -- adi.jl:97-128."""
    alg, abstol, observer = cache.alg, cache.abstol, cache.observer
    with _timeit("shifts"):
        mu = cache.shifts_oracle.take()
    cache.shifts.append(complex(mu))
    _observe(observer, "observe_gale_metadata", "ADI shifts", mu)
    if np.imag(mu) == 0:
        with _timeit("solve (real)"):      # (the fused C call also holds the residual update of adi.jl:171)
            perform_single_step_(cache, float(np.real(mu)))
    else:
        with _timeit("solve (complex)"):
            perform_double_step_(cache, complex(mu))
    want_compress = alg.compression and cache.last_compression >= alg.compression_interval
    on_lane = want_compress and ASYNC_COMPRESS and _fused_inner(alg)   # (every rank of a sharded run has its own lane)
    if want_compress and not on_lane:
        compress_cache_(cache)
    res_norm = cache.residual_norm = _dist.agree_scalar(_residual_norm_overlapped(cache))
    if on_lane:
        _start_async_compress(cache)   # after the norm: its synchronisation guarantees the terms are complete
    i = len(cache.shifts)
    _observe(observer, "observe_gale_step", i, cache.X, cache.residual, res_norm)
    if res_norm <= abstol:
        return
    if i < alg.maxiters:
        return
    _observe(observer, "observe_gale_failed")
    if alg.warn_convergence:
        warnings.warn(f"ADI did not converge: residual={res_norm} abstol={abstol} maxiters={alg.maxiters}")


def solve_(cache: ADICache) -> LDLt:
    """CommonSolve.solve!(::ADICache) -- adi.jl:71-89."""
    while not isdone(cache):
        step_(cache)
    if cache.alg.compression and cache.last_compression > 0:
        compress_cache_(cache)
    _join_pending(cache.X)
    iters = len(cache.shifts)
    _observe(cache.observer, "observe_gale_done", iters, cache.X, cache.residual, cache.residual_norm)
    return cache.X


# =============================================================================================
# low-rank (F)GMRES (src/lyapunov/gmres.jl, types.jl:44-52) -- SURVEY 8f rank 2
# =============================================================================================
class GMRES:  # types.jl:44-52
    def __init__(self, *, maxiters=3, maxrestarts=0, reltol=None, abstol=None, ignore_initial_guess=False,
                 compression=True, preconditioner=None):
        self.maxiters, self.maxrestarts, self.reltol, self.abstol = maxiters, maxrestarts, reltol, abstol
        self.ignore_initial_guess, self.compression = ignore_initial_guess, compression
        self.preconditioner = preconditioner


def lyapunov_operator(E, A, X: LDLt) -> LDLt:
    """gmres.jl:105-117 -- L*X = A'XE + E'XA = a * lowrank([E'Z A'Z], [0 Y; Y 0]); two SpMMs into column views."""
    a, Z, Y = X.destructure()
    Y = np.asarray(Y)
    k = Z.ncols
    be = backend()
    Z2 = DeviceMatrix.empty(2 * k)
    spmm("E", Z, 1.0, Z2.cols(0, k), 0.0)
    AtZ = _adj_matmul(A, Z)
    be.check(be.lib.dre_mat_copy(be.h, Z2.cols(k, 2 * k).view, AtZ.view))
    Y2 = np.zeros((2 * k, 2 * k), order="F")
    Y2[:k, k:] = Y
    Y2[k:, :k] = Y
    return LDLt([a], [Z2], [Y2])


def specialize(alg, prob):
    """gmres.jl:119-134 -- shift parameters that depend only on (E, A) are computed once per GMRES solve."""
    if isinstance(alg, Cyclic):
        return Cyclic(specialize(alg.inner, prob))
    if isinstance(alg, Heuristic):
        return list(_take_many(shifts_init(alg, prob)))
    if isinstance(alg, ADI):
        out = ADI.__new__(ADI)
        out.__dict__.update(alg.__dict__)
        out.shifts = specialize(alg.shifts, prob)
        return out
    if isinstance(alg, GMRES):
        out = GMRES.__new__(GMRES)
        out.__dict__.update(alg.__dict__)
        out.preconditioner = specialize(alg.preconditioner, prob)
        return out
    return alg


def _solve_gmres(prob: GALEProblem, alg: GMRES, *, initial_guess=None, abstol=None, observer=None) -> LDLt:
    """CommonSolve.solve(::GALEProblem, ::GMRES) -- gmres.jl:7-103 (FGMRES, Algorithm 2.2 of Saad 1993, on
    low-rank iterates): residuals, operator applications, inner products, norms and compressions run on the GPU
    through the same C-ABI calls as the ADI path; the (maxiters+1) x maxiters Hessenberg problem on the host."""
    _observe(observer, "observe_gale_start", prob, alg)
    prob = _device_problem(prob)
    E, A, Cm = prob.E, prob.A, prob.C
    maxiters, maxrestarts, compression = alg.maxiters, alg.maxrestarts, alg.compression
    if alg.ignore_initial_guess or initial_guess is None:
        initial_guess = Cm.zero()
    X = initial_guess.to_device_()
    reltol = alg.reltol if alg.reltol is not None else prob.A.shape[0] * EPS
    if abstol is None:
        abstol = alg.abstol if alg.abstol is not None else reltol * norm(Cm)
    preconditioner = specialize(alg.preconditioner, prob)
    H = np.zeros((maxiters + 1, maxiters))
    b = np.zeros(maxiters + 1)
    m, residual_norm, restarts = 0, np.inf, 0
    for restarts in range(maxrestarts + 1):
        m = 0
        R0 = residual(prob, X)
        beta = residual_norm = norm(R0)
        _observe(observer, "observe_gale_step", 0, X, R0, beta)
        if beta <= abstol:
            break
        V = [None] * (maxiters + 1)
        Z = [None] * maxiters
        V[0] = (1.0 / beta) * R0
        b[:] = 0.0
        b[0] = beta
        y = np.zeros(0)
        for j in range(maxiters):
            if preconditioner is None:
                Z[j] = V[j]
            else:
                Z[j] = solve(GALEProblem(E, A, V[j]), preconditioner, observer=observer)
            W = lyapunov_operator(E, A, Z[j])
            if compression:
                compress_(W)
            for i in range(j + 1):
                H[i, j] = dot(V[i], W)
                W = W - H[i, j] * V[i]
            H[j + 1, j] = norm(W)
            V[j + 1] = (1.0 / H[j + 1, j]) * W
            m = j + 1
            Hm, bm = H[:m + 1, :m], b[:m + 1]
            y = np.linalg.lstsq(Hm, bm, rcond=None)[0]
            residual_norm = float(np.linalg.norm(bm - Hm @ y))
            if residual_norm <= abstol:
                break
            _observe(observer, "observe_gale_step", m, None, None, residual_norm)
            if compression:
                compress_(V[j + 1])
        for j in range(m):
            X = X + (-y[j]) * Z[j]
        if compression and X.rank() > 0:
            compress_(X)
        _observe(observer, "observe_gale_step", m, X, None, residual_norm)
        if residual_norm <= abstol:
            break
    if residual_norm > abstol:
        _observe(observer, "observe_gale_failed")
        warnings.warn("GMRES did not converge")
    iters = restarts * maxiters + m
    _observe(observer, "observe_gale_done", iters, X, None, residual_norm)
    return X


# =============================================================================================
# Riccati drivers (src/riccati/*.jl)
# =============================================================================================
class GDREProblem:  # riccati/types.jl:11-20
    def __init__(self, E, A, B, C, X0, tspan):
        self.E, self.A, self.B, self.C, self.X0, self.tspan = E, A, B, C, X0, tspan


class DRESolution:  # riccati/types.jl:35-39
    def __init__(self, X, K, t):
        self.X, self.K, self.t = X, K, t


class GAREProblem:  # riccati/types.jl:46-51
    def __init__(self, E, A, G, Q):
        self.E, self.A, self.G, self.Q = E, A, G, Q


class Ros1:  # DifferentialRiccatiEquations.jl:55-57
    def __init__(self, inner_alg=None):
        self.inner_alg = inner_alg


class Ros2:  # DifferentialRiccatiEquations.jl:58-60
    def __init__(self, inner_alg=None):
        self.inner_alg = inner_alg


def quadratic_forcing(_i, residual_norm):  # newton.jl:165
    return min(0.1, 0.9 * residual_norm)


def superlinear_forcing(i, _r):  # newton.jl:156
    return 1.0 / (i ** 3 + 1)


class Newton:  # riccati/types.jl:95-106
    def __init__(self, inner_alg=None, *, maxiters=5, reltol=None, abstol=None, inexact=True, inexact_hybrid=True,
                 inexact_forcing=quadratic_forcing, linesearch=True):
        self.inner_alg = inner_alg if inner_alg is not None else ADI()
        self.maxiters, self.reltol, self.abstol = maxiters, reltol, abstol
        self.inexact, self.inexact_hybrid = inexact, inexact_hybrid
        self.inexact_forcing, self.linesearch = inexact_forcing, linesearch


def _tstops(tspan, dt):
    t0, tf = tspan
    nsteps = int(math.floor((tf - t0) / dt + 1e-12))
    return [t0 + i * dt for i in range(nsteps + 1)]


def _feedback(Bd: DeviceMatrix, X: LDLt):
    """lowrank_ros1.jl:25-28 / 53-56: K = (B'L D alpha)(L'E); returns the n x k panel E'L too."""
    alpha, L, D = X.destructure()
    BtLD = gemm_tn(Bd, L) @ D
    if alpha != 1:
        BtLD = BtLD * alpha
    EtL = spmm("E", L)
    Kt = gemm_nn(EtL, BtLD.T)  # K' = E'L (B'LD)'
    return alpha, L, D, BtLD, EtL, Kt


def _upload_gdre(prob: GDREProblem):
    be = backend()
    be.ensure_pencil(prob.E, prob.A)
    Bd = _as_device(np.asarray(prob.B))
    Ctd = _as_device(np.asarray(prob.C).T)
    prob.X0.to_device_()
    return be, Bd, Ctd


def _solve_ros1(prob: GDREProblem, alg: Ros1, *, dt, save_state, observer) -> DRESolution:
    """src/riccati/lowrank_ros1.jl:3-66."""
    _observe(observer, "observe_gdre_start", prob, alg)
    be, Bd, Ctd = _upload_gdre(prob)
    E, A = PencilCombo(0.0, 1.0), PencilCombo(1.0, 0.0)
    q = Ctd.ncols
    X = prob.X0
    tstops = _tstops(prob.tspan, dt)
    Xs = [X]
    alpha, L, D, BtLD, EtL, Kt = _feedback(Bd, X)
    Ks = [Kt.to_host().T.copy()]
    _observe(observer, "observe_gdre_step", tstops[0], X, Ks[-1])
    inner_alg = alg.inner_alg if alg.inner_alg is not None else ADI()
    for i in range(1, len(tstops)):
        tau = tstops[i - 1] - tstops[i]
        F = LowRankUpdate(A - E / (2 * tau), -1.0, Bd, Kt)  # lr_update(A - E/(2tau), -1, B, K), :39
        G = hcat([Ctd, EtL])  # :42
        S = _dcat([np.eye(q), BtLD.T @ BtLD + D / tau])  # :43
        R = compress_(LDLt([1.0], [G], [S]))  # :44
        lyap = GALEProblem(E, F, R)
        X = solve(lyap, inner_alg, observer=observer, initial_guess=X)  # :47-49
        if save_state:
            Xs.append(X)
        alpha, L, D, BtLD, EtL, Kt = _feedback(Bd, X)
        Ks.append(Kt.to_host().T.copy())
        _observe(observer, "observe_gdre_step", tstops[i], X, Ks[-1])
    if not save_state:
        Xs.append(X)
    _observe(observer, "observe_gdre_done")
    return DRESolution(Xs, Ks, tstops)


def _solve_ros2(prob: GDREProblem, alg: Ros2, *, dt, save_state, observer) -> DRESolution:
    """src/riccati/lowrank_ros2.jl:3-89."""
    _observe(observer, "observe_gdre_start", prob, alg)
    be, Bd, Ctd = _upload_gdre(prob)
    E, A = PencilCombo(0.0, 1.0), PencilCombo(1.0, 0.0)
    q = Ctd.ncols
    X = prob.X0
    tstops = _tstops(prob.tspan, dt)
    gamma = 1 + 1 / math.sqrt(2)
    Xs = [X]
    alpha, L, D, BtLD, EtL, Kt = _feedback(Bd, X)
    Ks = [Kt.to_host().T.copy()]
    _observe(observer, "observe_gdre_step", tstops[0], X, Ks[-1])
    inner_alg = alg.inner_alg if alg.inner_alg is not None else ADI()
    for i in range(1, len(tstops)):
        tau = tstops[i - 1] - tstops[i]
        gt = gamma * tau
        F = LowRankUpdate(gt * A - E / 2, 1.0 / (-gt), Bd, Kt)  # :41
        AtL = spmm("A", L)
        G = hcat([Ctd, AtL, EtL])  # :44
        n_G, n_L = G.ncols, L.ncols
        S = np.zeros((n_G, n_G), order="F")
        S[:q, :q] = np.eye(q)
        S[q:q + n_L, n_G - n_L:] = D
        S[n_G - n_L:, q:q + n_L] = D
        S[n_G - n_L:, n_G - n_L:] = -(BtLD.T @ BtLD)
        R1 = compress_(LDLt([1.0], [G], [S]))
        K1 = solve(GALEProblem(E, F, R1), inner_alg, observer=observer)  # :57-58
        kappa, T1, D1 = K1.destructure()
        BtT1D1 = gemm_tn(Bd, T1) @ D1
        if kappa != 1:
            BtT1D1 = BtT1D1 * kappa
        G2 = spmm("E", T1)
        S2 = (tau ** 2 * BtT1D1).T @ BtT1D1 + (2 - 1 / gamma) * D1
        R2 = LDLt([1.0], [G2], [np.asfortranarray(S2)])
        K2 = solve(GALEProblem(E, F, R2), inner_alg, observer=observer)  # :68-69
        X = X + ((2 - 1 / (2 * gamma)) * tau) * K1 + (-tau / 2) * K2  # :72
        if save_state:
            Xs.append(X)
        alpha, L, D, BtLD, EtL, Kt = _feedback(Bd, X)
        Ks.append(Kt.to_host().T.copy())
        _observe(observer, "observe_gdre_step", tstops[i], X, Ks[-1])
    if not save_state:
        Xs.append(X)
    _observe(observer, "observe_gdre_done")
    return DRESolution(Xs, Ks, tstops)


@_timed("residual(::GAREProblem, ::LDLt)")
def gare_residual(prob: GAREProblem, X: LDLt, *, AtL=None, EtL=None, BtLD=None, DLtGLD=None, _dev=None) -> LDLt:
    """src/riccati/residual.jl:6-52."""
    be = backend()
    Bd, Ctd = _dev
    Q, G = prob.Q, prob.G
    if X.iszero():
        return LDLt(list(Q.alphas), [L.copy() for L in Q.Ls], [np.array(D, order="F") for D in Q.Ds])
    gamma, _Ct, S = Q.destructure()
    beta, _B, Rinv = G.destructure()
    alpha, L, D = X.destructure()
    h, zk = Ctd.ncols, L.ncols
    dim = h + 2 * zk
    AtL = AtL if AtL is not None else spmm("A", L)
    EtL = EtL if EtL is not None else spmm("E", L)
    if DLtGLD is None:
        if BtLD is None:
            BtLD = gemm_tn(Bd, L) @ D
            if alpha * beta != 1:
                BtLD = BtLD * (alpha * beta)
        DLtGLD = BtLD.T @ Rinv @ BtLD
    R = hcat([Ctd, AtL, EtL])
    T = np.zeros((dim, dim), order="F")
    T[:h, :h] = gamma * S
    T[h:h + zk, h + zk:] = alpha * D
    T[h + zk:, h:h + zk] = T[h:h + zk, h + zk:]
    T[h + zk:, h + zk:] = -DLtGLD
    return compress_(LDLt([1.0], [R], [T]))


def _solve_newton(prob: GAREProblem, alg: Newton, *, observer=None) -> LDLt:
    """src/riccati/newton.jl:3-147."""
    _observe(observer, "observe_gare_start", prob, alg)
    be = backend()
    be.ensure_pencil(prob.E, prob.A)
    prob.G.to_device_()
    prob.Q.to_device_()
    a0, Bm, _ = prob.G.destructure()
    assert a0 == 1, "Scaled prob.G not yet implemented"
    a0, Ctm, _ = prob.Q.destructure()
    assert a0 == 1, "Scaled prob.Q not yet implemented"
    Bd, Ctd = Bm, Ctm  # device panels (lowrank() uploaded them)
    E, A = PencilCombo(0.0, 1.0), PencilCombo(1.0, 0.0)
    res = prob.Q
    res_norm = norm(res)
    n = be.n
    reltol = alg.reltol if alg.reltol is not None else n * EPS
    abstol = alg.abstol if alg.abstol is not None else reltol * res_norm
    X = LDLt([1.0], [DeviceMatrix.empty(0)], [np.zeros((0, 0))])
    i = 0
    X_prev = None
    inner_alg = alg.inner_alg
    inner_reltol = inner_alg.reltol if getattr(inner_alg, "reltol", None) is not None else reltol / 10
    m, q = Bd.ncols, Ctd.ncols

    def aux(Xc):
        alpha, L, D = Xc.destructure()
        EtL = spmm("E", L) if L.ncols else DeviceMatrix.empty(0)
        BtLD = gemm_tn(Bd, L) @ D if L.ncols else np.zeros((m, 0))
        if alpha != 1:
            BtLD = BtLD * alpha
        DLtGLD = BtLD.T @ BtLD
        Kt = gemm_nn(EtL, BtLD.T) if L.ncols else None
        return EtL, BtLD, DLtGLD, Kt

    while True:
        EtL, BtLD, DLtGLD, Kt = aux(X)
        res = gare_residual(prob, X, EtL=EtL, DLtGLD=DLtGLD, _dev=(Bd, Ctd))
        res_norm_prev = res_norm
        res_norm = norm(res)
        if i > 0 and alg.linesearch:
            a_ = 0.1
            if res_norm > (1 - a_) * res_norm_prev:
                X_tilde = X
                beta_ = 0.5
                lam = beta_
                while True:
                    X = (1 - lam) * X_prev + lam * X_tilde
                    res = gare_residual(prob, X, _dev=(Bd, Ctd))
                    res_norm = norm(res)
                    if res_norm < (1 - lam * a_) * res_norm_prev:
                        EtL, BtLD, DLtGLD, Kt = aux(X)
                        break
                    lam *= beta_
                    if lam < EPS:
                        warnings.warn("Line search failed; using un-modified iterate")
                        lam = 1.0
                        X = X_tilde
                        break
                _observe(observer, "observe_gare_metadata", "line search", lam)
        _observe(observer, "observe_gare_step", i, X, res, res_norm)
        if res_norm <= abstol:
            break
        if i >= alg.maxiters:
            _observe(observer, "observe_gare_failed")
            warnings.warn("Newton method did not converge")
            break
        i += 1
        if Kt is None:  # X == 0: K = 0
            Kt = DeviceMatrix.from_host(np.zeros((n, m)))
            EtXB = DeviceMatrix.from_host(np.zeros((n, m)))
        else:
            EtXB = gemm_nn(EtL, BtLD.T)  # E'L (B'LD)' = E'XB
        F = LowRankUpdate(A, -1.0, Bd, Kt)  # :103
        G = hcat([Ctd, EtXB])  # :106-111
        S = _dcat([np.eye(q), np.eye(m)])
        RHS = LDLt([1.0], [G], [S])
        lyap = GALEProblem(E, F, RHS)
        if alg.inexact:
            eta = alg.inexact_forcing(i, res_norm)
            inner_abstol = eta * res_norm
            if alg.inexact_hybrid:
                classical_abstol = inner_reltol * norm(lyap.C)
                switch_back = classical_abstol > inner_abstol
                _observe(observer, "observe_gare_metadata", "inexact", not switch_back)
                if switch_back:
                    inner_abstol = classical_abstol
            else:
                _observe(observer, "observe_gare_metadata", "inexact", True)
        else:
            inner_abstol = inner_reltol * norm(lyap.C)
        X_prev = X
        X = solve(lyap, inner_alg, abstol=inner_abstol, initial_guess=X_prev, observer=observer)
    _observe(observer, "observe_gare_done", i, X, res, res_norm)
    return X


def solve(prob, alg, **kw):
    """CommonSolve.solve for every problem/algorithm pair of the hot path
    (src/DifferentialRiccatiEquations.jl:78-94 for GDRE; solve = solve!(init(...)) for GALE)."""
    if isinstance(prob, GALEProblem) and isinstance(alg, ADI):
        return solve_(init(prob, alg, **kw))
    if isinstance(prob, GALEProblem) and isinstance(alg, GMRES):
        return _solve_gmres(prob, alg, **kw)
    if isinstance(prob, GDREProblem):
        dt = kw.pop("dt")
        save_state = kw.pop("save_state", False)
        observer = kw.pop("observer", None)
        if isinstance(alg, Ros1):
            return _solve_ros1(prob, alg, dt=dt, save_state=save_state, observer=observer)
        if isinstance(alg, Ros2):
            return _solve_ros2(prob, alg, dt=dt, save_state=save_state, observer=observer)
    if isinstance(prob, GAREProblem) and isinstance(alg, Newton):
        return _solve_newton(prob, alg, **kw)
    if isinstance(prob, BlockLinearProblem):
        return solve_block(prob, alg, **kw)
    raise TypeError(f"solve: unsupported combination {type(prob).__name__}, {type(alg).__name__}")


def upload_pencil(E, A):
    """Upload E, A (runs the symbolic analysis once).  Needed before ``lowrank`` can place factors on
    the device when building X0 / G / Q for a problem."""
    backend().ensure_pencil(E, A)
