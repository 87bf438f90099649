"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.

A NumPy/SciPy restatement of the low-rank (LRSIF, X = L D L^T) ADI hot path of
DifferentialRiccatiEquations.jl v0.5.5.  Every function cites the reference file:line it follows
(paths relative to /root/reference/).  It is imported only by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, as
the *checker* / CPU baseline -- the product path (``differentialriccatiequations.jl_b200``) never
imports it and fails loudly when the CUDA library is missing.

PARITY PINNING.  The reference is pure Julia; ``julia`` is not installed in this image or on the
GPU box, and the reference tree holds no golden vectors (all of its test inputs are unseeded
random).  The oracle is therefore pinned against every *known-answer / cross-check* the reference's
own tests hold for this path (see tests/test_oracle_pins.py):
  * test/Shifts.jl:165-183   Projection(2) on the Penzl 3x3 example -> the single shift -5/6
  * test/Shifts.jl:185-226   conjugate pairs stay adjacent (safe_sort!, Projection output)
  * test/runtests.jl:12-19   orth(zeros(4,1)) is 4x0
  * test/LDLt.jl:44-90       norm / compress! invariants
  * test/LowRankUpdate.jl:20-51  SMW solve satisfies M*X ~ B
  * test/residual.jl:18-29   low-rank residual norm == dense residual norm
  * test/tiny_random.jl:37-57  ADI vs dense Bartels-Stewart: delta < 1e-10, stepping API
  * test/rail.jl:52-70       low-rank Ros1/Ros2 K[end] == dense Ros1/Ros2 K[end]
  * test/rail.jl:74-88       Newton-ADI residual < 1e-10 ||Q||
Bit-level agreement with the Julia/UMFPACK/CHOLMOD/OpenBLAS stack is **unpinned** (those libraries
are third-party, un-vendored, and version-unpinned: Manifest.toml is git-ignored, only
``julia = "1.10"`` is fixed, Project.toml:30).  Sparse solves here use SciPy SuperLU, dense
factorizations use SciPy's LAPACK (geqp3, syevd/syevr, gesdd, ggev) -- the same LAPACK routines
Julia's LinearAlgebra calls.
"""
from __future__ import annotations

import copy
import itertools
import math
import time
import warnings
from contextlib import contextmanager

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

EPS = float(np.finfo(np.float64).eps)

# --------------------------------------------------------------------------------------------
# timing sections (same labels as the reference's @timeit_debug sections)
# --------------------------------------------------------------------------------------------
TIMERS: dict[str, float] = {}
COUNTS: dict[str, int] = {}


@contextmanager
def timeit(label: str):
    t0 = time.perf_counter()
    try:
        yield
    finally:
        TIMERS[label] = TIMERS.get(label, 0.0) + time.perf_counter() - t0
        COUNTS[label] = COUNTS.get(label, 0) + 1


def reset_timers():
    TIMERS.clear()
    COUNTS.clear()


def _observe(observer, name, *args):
    """src/Callbacks.jl:97-187 -- default methods are no-ops on ::Any."""
    if observer is None:
        return
    fn = getattr(observer, name, None)
    if fn is not None:
        with timeit("callbacks"):
            fn(*args)


# --------------------------------------------------------------------------------------------
# Stuff.jl
# --------------------------------------------------------------------------------------------
def restrict(A, Q):
    """src/Stuff.jl:9 (Q' A Q) and src/util/restrict.jl:5-8 (LowRankUpdate-aware)."""
    if isinstance(A, LowRankUpdate):
        return restrict(A.A, Q) + (1.0 / A.alpha) * ((Q.T @ A.U) @ (A.V @ Q))
    AQ = A @ Q
    return Q.T @ AQ


def orth(N):
    """src/Stuff.jl:13-18 -- SVD basis, keep singular values > n*eps (absolute)."""
    if sp.issparse(N):
        N = N.toarray()
    if N.shape[1] == 0:
        return np.zeros((N.shape[0], 0))
    U, s, _ = sla.svd(N, full_matrices=False, lapack_driver="gesdd")
    eps_ = N.shape[0] * EPS
    ids = np.nonzero(np.abs(s) > eps_)[0]
    return U[:, ids]


def delta(a, b):
    """src/Stuff.jl:21."""
    return np.linalg.norm(a - b) / max(np.linalg.norm(a), np.linalg.norm(b))


# --------------------------------------------------------------------------------------------
# LDLt.jl
# --------------------------------------------------------------------------------------------
def orthf(L):
    """src/LDLt.jl:237-245 -- pivoted QR, R un-permuted so that L ~ Q R."""
    with timeit("orthf"):
        Q, R, p = sla.qr(L, mode="economic", pivoting=True)
        ip = np.empty_like(p)
        ip[p] = np.arange(len(p))
        return Q, R[:, ip]


class LDLt:
    """src/LDLt.jl:29-33 -- lazy sum_i alpha_i L_i D_i L_i^T."""

    def __init__(self, alphas, Ls, Ds):
        self.alphas = list(alphas)
        self.Ls = list(Ls)
        self.Ds = list(Ds)

    # src/LDLt.jl:54-60 (destructuring; compresses iff more than one term)
    def destructure(self):
        if len(self.Ls) > 1:
            compress(self)
        assert len(self.alphas) == 1
        return self.alphas[0], self.Ls[0], self.Ds[0]

    def __iter__(self):
        return iter(self.destructure())

    @property
    def n(self):
        return self.Ls[0].shape[0]

    def rank(self):  # src/LDLt.jl:112
        return sum(L.shape[1] for L in self.Ls)

    def iszero(self):  # src/LDLt.jl:114
        return all(a == 0 for a in self.alphas) or self.rank() == 0

    def zero(self):  # src/LDLt.jl:116-121
        return lowrank(np.zeros((self.n, 0)), np.zeros((0, 0)))

    def to_dense(self):  # src/LDLt.jl:42-51
        M = np.zeros((self.n, self.n))
        for a, L, D in zip(self.alphas, self.Ls, self.Ds):
            M += L @ (a * D) @ L.T
        return M

    def __add__(self, other):  # src/LDLt.jl:131-148
        if self.n != other.n:
            raise ValueError("outer dimensions must match")
        if self.iszero():
            return other
        if other.iszero():
            return self
        return LDLt(self.alphas + other.alphas, self.Ls + other.Ls, self.Ds + other.Ds)

    def __neg__(self):  # src/LDLt.jl:150-153
        return LDLt([-a for a in self.alphas], self.Ls, self.Ds)

    def __sub__(self, other):
        return self + (-other)

    def __rmul__(self, alpha):  # src/LDLt.jl:156-159
        return LDLt([alpha * a for a in self.alphas], self.Ls, self.Ds)

    def __eq__(self, other):  # src/LDLt.jl:35
        return (self.alphas == other.alphas and len(self.Ls) == len(other.Ls)
                and all(np.array_equal(a, b) for a, b in zip(self.Ls, other.Ls))
                and all(np.array_equal(a, b) for a, b in zip(self.Ds, other.Ds)))


def lowrank(L, D=None):
    """src/LDLt.jl:24-27.  ``D=None`` stands for the UniformScaling ``I`` and is expanded."""
    L = np.asarray(L, dtype=float)
    if D is None:
        D = np.eye(L.shape[1])
    return LDLt([1.0], [L], [np.asarray(D, dtype=float)])


def _hcat(Xs):
    """src/util/_hcat.jl:5-18."""
    return np.concatenate(list(Xs), axis=1)


def _dcat(Xs, alphas=None):
    """src/util/_dcat.jl:8-22 (per-block alpha scaling)."""
    Xs = list(Xs)
    alphas = list(alphas) if alphas is not None else [1.0] * len(Xs)
    n = sum(X.shape[0] for X in Xs)
    D = np.zeros((n, n))
    k = 0
    for X, a in zip(Xs, alphas):
        l = X.shape[0]
        D[k:k + l, k:k + l] = X * a
        k += l
    return D


def concatenate(X: LDLt) -> LDLt:
    """src/LDLt.jl:174-191 (in place)."""
    if len(X.alphas) == 1:
        return X
    with timeit("concatenate!"):
        L = _hcat(X.Ls)
        D = _dcat(X.Ds, X.alphas)
        X.alphas[:] = [1.0]
        X.Ls[:] = [L]
        X.Ds[:] = [D]
    return X


def compress(X: LDLt) -> LDLt:
    """src/LDLt.jl:204-225 (in place)."""
    with timeit("compress!"):
        concatenate(X)
        L = X.Ls[0]
        D = X.Ds[0]
        Q, R = orthf(L)
        S = R @ D @ R.T
        S = 0.5 * (S + S.T) if False else S  # Symmetric(S) reads the upper triangle only
        with timeit("eigen"):
            lam, V = sla.eigh(S, lower=False)
        eps_ = 100 * np.max(np.abs(lam)) * EPS
        ids = np.nonzero(np.abs(lam) >= eps_)[0]
        X.Ls[0] = Q @ V[:, ids]
        X.Ds[0] = np.diag(lam[ids])
    return X


def norm(X: LDLt) -> float:
    """src/LDLt.jl:77-89."""
    with timeit("norm(::LDLt)"):
        concatenate(X)
        a, L, D = X.alphas[0], X.Ls[0], X.Ds[0]
        if L.shape[1] == 0:
            return 0.0
        _, R = orthf(L)
        return abs(a) * float(np.linalg.norm(R @ D @ R.T))


# --------------------------------------------------------------------------------------------
# LowRankUpdate.jl, blocklinear/*.jl
# --------------------------------------------------------------------------------------------
class LowRankUpdate:
    """src/LowRankUpdate.jl:18-26 -- lazy A + inv(alpha) U V."""

    def __init__(self, A, alpha, U, V):
        self.A, self.alpha, self.U, self.V = A, alpha, U, V

    @property
    def shape(self):
        return self.A.shape

    def adjoint(self):  # :51-54
        At = self.A.T if not isinstance(self.A, Factorized) else self.A.adjoint()
        return LowRankUpdate(At, np.conj(self.alpha), self.V.conj().T, self.U.conj().T)

    def plus_sparse(self, E):  # :66-70
        assert sp.issparse(E)
        return LowRankUpdate((self.A + E).tocsc(), self.alpha, self.U, self.V)

    def matmul(self, X):  # :72-86
        if X.ndim == 2 and X.shape[0] == X.shape[1]:
            warnings.warn("Multiplying LowRankUpdate by square matrix; memory usage may increase severely")
        Y = self.A @ X
        Y = Y + (1.0 / self.alpha) * (self.U @ (self.V @ X))
        return Y

    def __matmul__(self, X):
        return self.matmul(X)

    def to_dense(self):  # :56-59
        A = self.A.toarray() if sp.issparse(self.A) else self.A
        return A + (1.0 / self.alpha) * (self.U @ self.V)


def lr_update(A, alpha, U, V):
    """src/LowRankUpdate.jl:38-39."""
    if sp.issparse(A):
        return LowRankUpdate(A, alpha, U, V)
    return A + ((1.0 / alpha) * U) @ V


def adjoint(A):
    if isinstance(A, LowRankUpdate):
        return A.adjoint()
    return A.T.tocsc() if sp.issparse(A) else A.conj().T


class Factorized:
    """Result of ``factorize`` (LinearAlgebra / SuiteSparse in the reference, SuperLU / LAPACK here)."""

    def __init__(self, A):
        if sp.issparse(A):
            with timeit("factorize (sparse)"):
                self.kind = "sparse"
                self.lu = spla.splu(A.tocsc(), permc_spec="MMD_AT_PLUS_A")
        else:
            self.kind = "dense"
            self.lu = sla.lu_factor(A)
        self.shape = A.shape

    def solve(self, B):
        if self.kind == "sparse":
            with timeit("solve (sparse)"):
                if np.iscomplexobj(self.lu.L.data) and not np.iscomplexobj(B):
                    B = B.astype(complex)
                return self.lu.solve(np.ascontiguousarray(B))
        return sla.lu_solve(self.lu, B)


def factorize(A):
    """src/blocklinear/types.jl:41-42 and src/LowRankUpdate.jl:88-91."""
    if isinstance(A, Factorized):
        return A
    if isinstance(A, LowRankUpdate):
        return LowRankUpdate(factorize(A.A), A.alpha, A.U, A.V)
    return Factorized(A)


def backslash(F, B):
    """``F \\ B`` for a factorization or a factorized LowRankUpdate (src/LowRankUpdate.jl:61-64)."""
    if isinstance(F, LowRankUpdate):
        return solve_block(BlockLinearProblem(F, B), ShermanMorrisonWoodbury())
    return F.solve(B)


class BlockLinearProblem:  # src/blocklinear/types.jl:10-13
    def __init__(self, A, B):
        self.A, self.B = A, B


class Backslash:  # src/blocklinear/types.jl:31-34
    def __init__(self, factorize_fn=None):
        self.factorize = factorize_fn if factorize_fn is not None else factorize


class ShermanMorrisonWoodbury:  # src/blocklinear/types.jl:35-39
    def __init__(self, ALG=None, alg=None):
        self.ALG = ALG if ALG is not None else Backslash()
        self.alg = alg if alg is not None else Backslash()


class BackslashSolver:  # src/blocklinear/backslash.jl:3-21
    def __init__(self, F, B):
        self.F, self.B = F, B

    def solve(self):
        return backslash(self.F, self.B)


class SMWSolver:  # src/blocklinear/sherman-morrison-woodbury.jl:3-45
    def __init__(self, prob, smw):
        with timeit("Sherman-Morrison-Woodbury"):
            if not isinstance(prob.A, LowRankUpdate):
                raise NotImplementedError("Not implemented")
            A, alpha, U, V = prob.A.A, prob.A.alpha, prob.A.U, prob.A.V
            B = prob.B
            self.AinvU = solve_block(BlockLinearProblem(A, U), smw.ALG)  # :19
            S = alpha * np.eye(V.shape[0]) + V @ self.AinvU  # :20
            self.V = V
            self.SOLVER = init_block(BlockLinearProblem(A, B), smw.ALG)  # :28
            self.S_fact = smw.alg.factorize(S)  # :29 (init of the inner dense solver)

    @property
    def B(self):
        return self.SOLVER.B

    def solve(self):
        with timeit("Sherman-Morrison-Woodbury"):
            AinvB = self.SOLVER.solve()  # :36
            T = self.V @ AinvB  # :38
            Y = backslash(self.S_fact, T)  # :39
            return AinvB - self.AinvU @ Y  # :43


def init_block(prob, alg):
    if isinstance(alg, Backslash):
        return BackslashSolver(alg.factorize(prob.A), prob.B)
    if isinstance(alg, ShermanMorrisonWoodbury):
        return SMWSolver(prob, alg)
    raise TypeError(alg)


def solve_block(prob, alg):
    return init_block(prob, alg).solve()


# --------------------------------------------------------------------------------------------
# Shifts.jl, shifts/*.jl
# --------------------------------------------------------------------------------------------
def safe_sort(shifts):
    """src/shifts/helpers.jl:122 -- sort by (real, |imag|), stable."""
    shifts = list(shifts)
    shifts.sort(key=lambda v: (np.real(v), abs(np.imag(v))))
    return shifts


def stabilize_ritz_values(lam, desc):
    """src/shifts/helpers.jl:129-140."""
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


class Strategy:
    pass


class Projection(Strategy):  # src/shifts/projection.jl:25-32
    def __init__(self, u):
        if u % 2 == 1:
            raise ValueError(f"History must be even; got {u}")
        self.n_history = u


class ProjectionShiftIterator:  # src/shifts/projection.jl:34-73
    def __init__(self, prob, n_history):
        self.prob, self.n_history, self.Vs = prob, n_history, []

    def update(self, X, R, *Vs):  # :45-52
        if not Vs:
            self.Vs.append(R)  # aliases R (mutated in place later, adi.jl:171)
        self.Vs.extend(Vs)
        lst = len(self.Vs)
        fst = max(0, lst - self.n_history)
        self.Vs = self.Vs[fst:lst]

    def take_many(self):  # :54-73
        E, A = self.prob.E, self.prob.A
        N = np.concatenate(self.Vs, axis=1)
        Q = orth(N)
        Et = restrict(E, Q)
        At = restrict(A, Q)
        lam = sla.eigvals(At, Et)
        if np.all(np.imag(lam) == 0):
            lam = np.real(lam)
        lam = stabilize_ritz_values(lam, "(A, E)")
        return safe_sort(lam)


class BufferedIterator:  # src/shifts/helpers.jl:70-75, 100-113
    def __init__(self, gen):
        self.buffer, self.generator = [], gen

    def update(self, *args):
        upd = getattr(self.generator, "update", None)
        if upd is not None:
            upd(*args)

    def take(self):
        if not self.buffer:
            self.buffer = list(self.generator.take_many())
        return self.buffer.pop(0)


class WrappedIterator:  # src/shifts/helpers.jl:85-104
    def __init__(self, func, gen):
        self.func, self.generator = func, gen

    def update(self, *args):
        upd = getattr(self.generator, "update", None)
        if upd is not None:
            upd(*args)

    def take_many(self):
        return self.func(_take_many(self.generator))

    def take(self):  # default take! = popfirst! (Shifts.jl:116) is not defined for this type
        raise TypeError("WrappedIterator must be wrapped in a BufferedIterator or Cyclic")


class _CycleIterator:  # Stateful(cycle(values)), helpers.jl:93
    def __init__(self, values):
        self.values = list(values)
        self._it = itertools.cycle(self.values)

    def update(self, *args):
        return None

    def take(self):
        return next(self._it)


class _ListIterator:  # plain vector: take! = popfirst! (Shifts.jl:116)
    def __init__(self, values):
        self.values = list(values)

    def update(self, *args):
        return None

    def take(self):
        return self.values.pop(0)

    def take_many(self):
        return self.values


class Cyclic(Strategy):  # src/shifts/helpers.jl:19-27
    def __init__(self, inner):
        self.inner = inner


class Wrapped(Strategy):  # src/shifts/helpers.jl:48-58
    def __init__(self, func, inner):
        self.func, self.inner = func, inner


class Heuristic(Strategy):  # src/shifts/heuristic.jl:22-30
    def __init__(self, nshifts, k_plus, k_minus, alg_E=None, alg_A=None):
        self.nshifts, self.k_plus, self.k_minus = nshifts, k_plus, k_minus
        self.alg_E = alg_E if alg_E is not None else Backslash()
        self.alg_A = alg_A if alg_A is not None else Backslash()


def _take_many(gen):
    if hasattr(gen, "take_many"):
        return gen.take_many()
    return list(gen)  # take_many!(values) = values (helpers.jl:103)


def heuristic(R, nshifts=None):
    """src/shifts/heuristic.jl:82-101 -- Penzl's greedy min-max selection."""
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


def compute_ritz_values(op, b0, k, desc):
    """src/shifts/heuristic.jl:103-130 -- Arnoldi with twice-repeated MGS."""
    n = len(b0)
    H = np.zeros((k + 1, k))
    V = np.zeros((n, k + 1))
    V[:, 0] = (1.0 / np.linalg.norm(b0)) * b0
    for j in range(k):
        w = np.array(op(V[:, j]), dtype=float).reshape(n)
        for _ in range(2):
            for i in range(j + 1):
                g = float(V[:, i] @ w)
                H[i, j] += g
                w -= V[:, i] * g
        beta = float(np.linalg.norm(w))
        H[j + 1, j] = beta
        V[:, j + 1] = (1.0 / beta) * w
    ritz = sla.eigvals(H[:k, :k])
    if np.all(np.imag(ritz) == 0):
        ritz = np.real(ritz)
    return stabilize_ritz_values(ritz, desc)


def _matmul(A, x):
    return A.matmul(x) if isinstance(A, LowRankUpdate) else A @ x


def shifts_init(strategy, prob):
    """Shifts.init for every strategy (projection.jl:40-43, heuristic.jl:39-66, helpers.jl:86-99)."""
    if isinstance(strategy, Projection):
        return BufferedIterator(ProjectionShiftIterator(prob, strategy.n_history))
    if isinstance(strategy, Heuristic):
        E, A = prob.E, prob.A
        n = E.shape[1]
        b0 = np.ones(n)
        solver_E = init_block(BlockLinearProblem(E, np.empty(n)), strategy.alg_E)

        def op_plus(x):
            solver_E.B = _matmul(A, x)
            return _solve_solver(solver_E)

        R_plus = compute_ritz_values(op_plus, b0, strategy.k_plus, "E^-1 A")
        solver_A = init_block(BlockLinearProblem(A, np.empty(n)), strategy.alg_A)

        def op_minus(x):
            solver_A.B = _matmul(E, x)
            return _solve_solver(solver_A)

        R_minus = compute_ritz_values(op_minus, b0, strategy.k_minus, "A^-1 E")
        R = list(R_plus) + [1.0 / v for v in R_minus]
        return _ListIterator(heuristic(R, strategy.nshifts))
    if isinstance(strategy, Cyclic):
        inner = strategy.inner
        vals = _take_many(shifts_init(inner, prob)) if isinstance(inner, Strategy) else list(inner)
        return _CycleIterator(vals)
    if isinstance(strategy, Wrapped):
        it = shifts_init(strategy.inner, prob)
        if isinstance(it, BufferedIterator):
            return BufferedIterator(WrappedIterator(strategy.func, it.generator))
        return WrappedIterator(strategy.func, it)
    # custom strategy protocol: objects with init(prob) -> iterator
    return strategy.init(prob)


def _solve_solver(solver):
    if isinstance(solver, BackslashSolver):
        return backslash(solver.F, solver.B)
    if isinstance(solver, SMWSolver):
        solver.SOLVER.B = solver.B if False else solver.SOLVER.B
        return solver.solve()
    raise TypeError(solver)


# --------------------------------------------------------------------------------------------
# lyapunov/types.jl, lyapunov/residual.jl, lyapunov/adi.jl
# --------------------------------------------------------------------------------------------
class GALEProblem:  # src/lyapunov/types.jl:10-16   A'XE + E'XA = -C
    def __init__(self, E, A, C):
        self.E, self.A, self.C = E, A, C


class ADI:  # src/lyapunov/types.jl:20-32
    def __init__(self, inner_alg=None, *, maxiters=100, reltol=None, abstol=None, shifts=None,
                 ignore_initial_guess=False, compression_interval=10, compression=True,
                 warn_convergence=True):
        self.maxiters, self.reltol, self.abstol = maxiters, reltol, abstol
        self.shifts = shifts if shifts is not None else Projection(2)
        self.ignore_initial_guess = ignore_initial_guess
        self.inner_alg = inner_alg if inner_alg is not None else Backslash()
        self.compression_interval, self.compression = compression_interval, compression
        self.warn_convergence = warn_convergence


def _adj_matmul(A, L):
    """A' * L for sparse A or LowRankUpdate A (LowRankUpdate.jl:51-54, 77-86)."""
    if isinstance(A, LowRankUpdate):
        return A.adjoint().matmul(L)
    return A.T @ L


def gale_residual(prob: GALEProblem, val: LDLt) -> LDLt:
    """src/lyapunov/residual.jl:3-31."""
    with timeit("residual(::GALEProblem, ::LDLt)"):
        E, A, C = prob.E, prob.A, prob.C
        if val.iszero():
            return copy.deepcopy(C)
        alpha, G, S = C.destructure()
        beta, L, D = val.destructure()
        n_G, n_0 = G.shape[1], L.shape[1]
        dim = n_G + 2 * n_0
        R = _hcat([G, E.T @ L, _adj_matmul(A, L)])
        T = np.zeros((dim, dim))
        T[:n_G, :n_G] = alpha * S
        T[n_G:n_G + n_0, n_G + n_0:] = beta * D
        T[n_G + n_0:, n_G:n_G + n_0] = T[n_G:n_G + n_0, n_G + n_0:]
        return compress(lowrank(R, T))


def gale_residual_dense(prob: GALEProblem, X: np.ndarray) -> np.ndarray:
    """src/lyapunov/residual.jl:33-42 (test only)."""
    E, A = prob.E, prob.A
    Ad = A.to_dense() if isinstance(A, LowRankUpdate) else (A.toarray() if sp.issparse(A) else A)
    Ed = E.toarray() if sp.issparse(E) else E
    C = prob.C.to_dense() if isinstance(prob.C, LDLt) else prob.C
    return C + Ad.T @ X @ Ed + Ed.T @ X @ Ad


class ADICache:
    """src/lyapunov/adi.jl:5-21."""

    def __init__(self, **kw):
        self.last_compression = 0
        self.V1 = self.V2 = None
        self.__dict__.update(kw)

    # src/lyapunov/adi.jl:91-95 (iteration protocol)
    def __iter__(self):
        done = False
        while not done:
            adi_step(self)
            done = adi_isdone(self)
            yield self


def adi_init(prob: GALEProblem, alg: ADI, *, initial_guess=None, initial_residual=None, abstol=None,
             observer=None) -> ADICache:
    """src/lyapunov/adi.jl:29-69."""
    _observe(observer, "observe_gale_start", prob, alg)
    E, A, C = prob.E, prob.A, prob.C
    if alg.ignore_initial_guess or initial_guess is None:
        initial_guess = C.zero()
    if initial_residual is None:
        initial_residual = gale_residual(prob, initial_guess)
    X = initial_guess
    _, R, _T = initial_residual.destructure()
    residual_norm = norm(initial_residual)
    with timeit("shifts"):
        oracle = shifts_init(alg.shifts, prob)
        oracle.update(X, R)
        shifts = []
    reltol = alg.reltol if alg.reltol is not None else A.shape[0] * EPS
    if abstol is None:
        abstol = alg.abstol if alg.abstol is not None else reltol * norm(C)
    _observe(observer, "observe_gale_step", 0, X, initial_residual, residual_norm)
    increment = initial_residual.zero()
    return ADICache(prob=prob, alg=alg, abstol=abstol, observer=observer, shifts_oracle=oracle,
                    shifts=shifts, X=X, increment=increment, residual=initial_residual,
                    residual_norm=residual_norm)


def adi_isdone(cache: ADICache) -> bool:
    """src/lyapunov/adi.jl:130-141."""
    if cache.residual_norm <= cache.abstol:
        return True
    niters = len(cache.shifts)
    if niters > 0 and cache.increment.iszero():
        return True
    return niters >= cache.alg.maxiters


def adi_compress(cache: ADICache):
    """src/lyapunov/adi.jl:143-147."""
    compress(cache.X)
    cache.last_compression = 0


def _shifted_operator(A, E, mu):
    """F = A' + (mu*E)'  (adi.jl:156) resp. A' + (conj(mu)*E)' = A' + mu E' (adi.jl:195)."""
    muEt = (mu * E.T).tocsc()
    if isinstance(A, LowRankUpdate):
        return A.adjoint().plus_sparse(muEt)
    return (A.T + muEt).tocsc()


def perform_single_step(cache: ADICache, mu: float):
    """src/lyapunov/adi.jl:149-179."""
    prob, alg, residual = cache.prob, cache.alg, cache.residual
    E, A = prob.E, prob.A
    alpha, R, T = residual.destructure()
    F = _shifted_operator(A, E, mu)
    with timeit("solve (real)"):
        V = solve_block(BlockLinearProblem(F, R), alg.inner_alg)
    increment = (-2 * mu * alpha) * lowrank(V, T)
    cache.increment = increment
    with timeit("spmm residual update"):
        R += (-2 * mu) * (E.T @ V)  # mul!(R, E', V, -2mu, true): in place
    cache.X = cache.X + cache.increment
    cache.last_compression += 1
    with timeit("shifts"):
        cache.shifts_oracle.update(cache.X, R, V)


def perform_double_step(cache: ADICache, mu: complex):
    """src/lyapunov/adi.jl:181-225."""
    prob, alg, residual = cache.prob, cache.alg, cache.residual
    E, A = prob.E, prob.A
    alpha, R, T = residual.destructure()
    with timeit("shifts"):
        mu_next = cache.shifts_oracle.take()
    assert np.isclose(mu_next, np.conj(mu)), (mu, mu_next)
    cache.shifts.append(complex(mu_next))
    _observe(cache.observer, "observe_gale_metadata", "ADI shifts", mu_next)
    F = _shifted_operator(A, E, mu)  # A' + (conj(mu) E)' == A' + mu E'
    with timeit("solve (complex)"):
        V = solve_block(BlockLinearProblem(F, R), alg.inner_alg)
    if not np.any(V):
        warnings.warn("Increment is zero")
        cache.increment = residual.zero()
        return
    d = mu.real / mu.imag
    Vr = np.real(V)
    Vi = np.imag(V)
    V1 = math.sqrt(2.0) * Vr + (math.sqrt(2.0) * d) * Vi
    V2 = math.sqrt(2 * d * d + 2) * Vi
    cache.increment = (-2 * mu.real * alpha) * (lowrank(V1, T) + lowrank(V2, T))
    with timeit("spmm residual update"):
        R += (-2 * math.sqrt(2.0) * mu.real) * (E.T @ V1)
    cache.X = cache.X + cache.increment
    cache.last_compression += 2
    with timeit("shifts"):
        cache.shifts_oracle.update(cache.X, R, V1, V2)


def adi_step(cache: ADICache):
    """src/lyapunov/adi.jl:97-128."""
    alg, abstol, observer = cache.alg, cache.abstol, cache.observer
    with timeit("shifts"):
        mu = cache.shifts_oracle.take()
    cache.shifts.append(complex(mu))
    _observe(observer, "observe_gale_metadata", "ADI shifts", mu)
    if np.imag(mu) == 0:
        perform_single_step(cache, float(np.real(mu)))
    else:
        perform_double_step(cache, complex(mu))
    if alg.compression and cache.last_compression >= alg.compression_interval:
        adi_compress(cache)
    res_norm = cache.residual_norm = norm(cache.residual)
    i = len(cache.shifts)
    _observe(observer, "observe_gale_step", i, cache.X, cache.residual, res_norm)
    if res_norm <= abstol:
        return
    if i < alg.maxiters:
        return
    _observe(observer, "observe_gale_failed")
    if alg.warn_convergence:
        warnings.warn(f"ADI did not converge: residual={res_norm} abstol={abstol} maxiters={alg.maxiters}")


def adi_solve(cache: ADICache) -> LDLt:
    """src/lyapunov/adi.jl:71-89."""
    while not adi_isdone(cache):
        adi_step(cache)
    if cache.alg.compression and cache.last_compression > 0:
        adi_compress(cache)
    iters = len(cache.shifts)
    _observe(cache.observer, "observe_gale_done", iters, cache.X, cache.residual, cache.residual_norm)
    return cache.X


def solve_gale(prob: GALEProblem, alg, **kw) -> LDLt:
    """CommonSolve.solve(prob, alg; kw...) = solve!(init(prob, alg; kw...)) for ADI; GMRES has its own driver."""
    if isinstance(alg, GMRES):
        return solve_gale_gmres(prob, alg, **kw)
    return adi_solve(adi_init(prob, alg, **kw))


# --------------------------------------------------------------------------------------------
# LDLt.jl dot, lyapunov/gmres.jl  (low-rank FGMRES with an ADI preconditioner)
# --------------------------------------------------------------------------------------------
def dot(X1: LDLt, X2: LDLt) -> float:
    """src/LDLt.jl:91-108 -- Frobenius inner product tr(X1' X2) without forming n x n matrices."""
    if X1.n != X2.n:
        raise ValueError("DimensionMismatch")
    concatenate(X1)
    concatenate(X2)
    alpha, A, B = X1.alphas[0], X1.Ls[0], X1.Ds[0]
    beta, C, D = X2.alphas[0], X2.Ls[0], X2.Ds[0]
    AtC = A.T @ C
    M = (B.T @ AtC @ D) * (alpha * beta)
    return float(np.sum(AtC * M))   # sum_i a_i' M c_i over the rows a_i of A, c_i of C


class GMRES:  # src/lyapunov/types.jl:44-52
    def __init__(self, *, maxiters=3, maxrestarts=0, reltol=None, abstol=None, ignore_initial_guess=False,
                 compression=True, preconditioner=None):
        self.maxiters, self.maxrestarts, self.reltol, self.abstol = maxiters, maxrestarts, reltol, abstol
        self.ignore_initial_guess, self.compression = ignore_initial_guess, compression
        self.preconditioner = preconditioner


def lyapunov_operator(E, A, X: LDLt) -> LDLt:
    """src/lyapunov/gmres.jl:105-117 -- L*X = A'XE + E'XA as a*lowrank([E'Z A'Z], [0 Y; Y 0])."""
    a, Z, Y = X.destructure()
    k = Y.shape[0]
    Z2 = _hcat([E.T @ Z, _adj_matmul(A, Z)])
    Y2 = np.zeros((2 * k, 2 * k))
    Y2[:k, k:] = Y
    Y2[k:, :k] = Y
    return a * lowrank(Z2, Y2)


def specialize(alg, prob):
    """src/lyapunov/gmres.jl:119-134 -- shift parameters that only depend on (E, A) are computed once."""
    if isinstance(alg, Cyclic):
        return Cyclic(specialize(alg.inner, prob))
    if isinstance(alg, Heuristic):
        return list(_take_many(shifts_init(alg, prob)))
    if isinstance(alg, (ADI, GMRES)):
        out = copy.copy(alg)
        if isinstance(alg, ADI):
            out.shifts = specialize(alg.shifts, prob)
        else:
            out.preconditioner = specialize(alg.preconditioner, prob)
        return out
    return alg


def solve_gale_gmres(prob: GALEProblem, alg: GMRES, *, initial_guess=None, abstol=None, observer=None) -> LDLt:
    """src/lyapunov/gmres.jl:7-103 (Algorithm 2.2 of Saad's FGMRES paper on low-rank iterates)."""
    _observe(observer, "observe_gale_start", prob, alg)
    E, A, C = prob.E, prob.A, prob.C
    maxiters, maxrestarts, compression = alg.maxiters, alg.maxrestarts, alg.compression
    if alg.ignore_initial_guess or initial_guess is None:
        initial_guess = C.zero()
    X = initial_guess
    reltol = alg.reltol if alg.reltol is not None else A.shape[0] * EPS
    if abstol is None:
        abstol = alg.abstol if alg.abstol is not None else reltol * norm(C)
    preconditioner = specialize(alg.preconditioner, prob)
    H = np.zeros((maxiters + 1, maxiters))
    b = np.zeros(maxiters + 1)
    m, residual_norm, restarts = 0, np.inf, 0
    for restarts in range(maxrestarts + 1):
        m = 0
        R0 = gale_residual(prob, X)
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
                Z[j] = solve_gale(GALEProblem(E, A, V[j]), preconditioner, observer=observer)
            W = lyapunov_operator(E, A, Z[j])
            if compression:
                compress(W)
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
                compress(V[j + 1])
        for j in range(m):
            X = X + (-y[j]) * Z[j]
        if compression:
            compress(X)
        _observe(observer, "observe_gale_step", m, X, None, residual_norm)
        if residual_norm <= abstol:
            break
    if residual_norm > abstol:
        _observe(observer, "observe_gale_failed")
        warnings.warn("GMRES did not converge")
    iters = restarts * maxiters + m
    _observe(observer, "observe_gale_done", iters, X, None, residual_norm)
    return X


# --------------------------------------------------------------------------------------------
# riccati/*.jl
# --------------------------------------------------------------------------------------------
class GDREProblem:  # src/riccati/types.jl:11-20
    def __init__(self, E, A, B, C, X0, tspan):
        self.E, self.A, self.B, self.C, self.X0, self.tspan = E, A, B, C, X0, tspan


class DRESolution:  # src/riccati/types.jl:35-39
    def __init__(self, X, K, t):
        self.X, self.K, self.t = X, K, t


class GAREProblem:  # src/riccati/types.jl:46-51   Q + A'XE + E'XA - E'XGXE = 0
    def __init__(self, E, A, G, Q):
        self.E, self.A, self.G, self.Q = E, A, G, Q


class Ros1:
    def __init__(self, inner_alg=None):
        self.inner_alg = inner_alg


class Ros2:
    def __init__(self, inner_alg=None):
        self.inner_alg = inner_alg


def _tstops(tspan, dt):
    """Julia range t0:dt:tf (lowrank_ros1.jl:19)."""
    t0, tf = tspan
    nsteps = int(math.floor((tf - t0) / dt + 1e-12))
    return [t0 + i * dt for i in range(nsteps + 1)]


def _feedback(E, B, X: LDLt):
    """lowrank_ros1.jl:25-28 / 53-56."""
    alpha, L, D = X.destructure()
    BtLD = (B.T @ L) @ D
    if alpha != 1:
        BtLD = BtLD * alpha
    K = BtLD @ (E.T @ L).T  # (L'E) = (E'L)'
    return alpha, L, D, BtLD, K


def solve_gdre_ros1(prob: GDREProblem, alg: Ros1, *, dt, save_state=False, observer=None) -> DRESolution:
    """src/riccati/lowrank_ros1.jl:3-66."""
    _observe(observer, "observe_gdre_start", prob, alg)
    E, A, B, C, tspan = prob.E, prob.A, prob.B, prob.C, prob.tspan
    q = C.shape[0]
    X = prob.X0
    tstops = _tstops(tspan, dt)
    Xs = [X]
    alpha, L, D, BtLD, K = _feedback(E, B, X)
    Ks = [K]
    _observe(observer, "observe_gdre_step", tstops[0], X, K)
    inner_alg = alg.inner_alg if alg.inner_alg is not None else ADI()
    for i in range(1, len(tstops)):
        tau = tstops[i - 1] - tstops[i]
        F = lr_update((A - E / (2 * tau)).tocsc(), -1.0, B, K)
        G = _hcat([C.T, E.T @ L])
        S = _dcat([np.eye(q), BtLD.T @ BtLD + D / tau])
        R = compress(lowrank(G, S))
        lyap = GALEProblem(E, F, R)
        with timeit("ADI"):
            X = solve_gale(lyap, inner_alg, observer=observer, initial_guess=X)
        if save_state:
            Xs.append(X)
        alpha, L, D, BtLD, K = _feedback(E, B, X)
        Ks.append(K)
        _observe(observer, "observe_gdre_step", tstops[i], X, K)
    if not save_state:
        Xs.append(X)
    _observe(observer, "observe_gdre_done")
    return DRESolution(Xs, Ks, tstops)


def solve_gdre_ros2(prob: GDREProblem, alg: Ros2, *, dt, save_state=False, observer=None) -> DRESolution:
    """src/riccati/lowrank_ros2.jl:3-89."""
    _observe(observer, "observe_gdre_start", prob, alg)
    E, A, B, C, tspan = prob.E, prob.A, prob.B, prob.C, prob.tspan
    q = C.shape[0]
    X = prob.X0
    tstops = _tstops(tspan, dt)
    gamma = 1 + 1 / math.sqrt(2)
    Xs = [X]
    alpha, L, D, BtLD, K = _feedback(E, B, X)
    Ks = [K]
    _observe(observer, "observe_gdre_step", tstops[0], X, K)
    inner_alg = alg.inner_alg if alg.inner_alg is not None else ADI()
    for i in range(1, len(tstops)):
        tau = tstops[i - 1] - tstops[i]
        gt = gamma * tau
        F = lr_update((gt * A - E / 2).tocsc(), 1.0 / (-gt), B, K)
        # stage 1 (:44-58)
        G = _hcat([C.T, A.T @ L, E.T @ L])
        n_G, n_L = G.shape[1], L.shape[1]
        S = np.zeros((n_G, n_G))
        b1 = slice(0, q)
        b2 = slice(q, q + n_L)
        b3 = slice(n_G - n_L, n_G)
        S[b1, b1] = np.eye(q)
        S[b2, b3] = D
        S[b3, b2] = D
        S[b3, b3] = -(BtLD.T @ BtLD)
        R1 = compress(lowrank(G, S))
        K1 = solve_gale(GALEProblem(E, F, R1), inner_alg, observer=observer)
        # stage 2 (:60-69)
        kappa, T1, D1 = K1.destructure()
        BtT1D1 = (B.T @ T1) @ D1
        if kappa != 1:
            BtT1D1 = BtT1D1 * kappa
        G2 = E.T @ T1
        S2 = (tau ** 2 * BtT1D1).T @ BtT1D1 + (2 - 1 / gamma) * D1
        R2 = lowrank(G2, S2)
        K2 = solve_gale(GALEProblem(E, F, R2), inner_alg, observer=observer)
        # update (:72)
        X = X + ((2 - 1 / (2 * gamma)) * tau) * K1 + (-tau / 2) * K2
        if save_state:
            Xs.append(X)
        alpha, L, D, BtLD, K = _feedback(E, B, X)
        Ks.append(K)
        _observe(observer, "observe_gdre_step", tstops[i], X, K)
    if not save_state:
        Xs.append(X)
    _observe(observer, "observe_gdre_done")
    return DRESolution(Xs, Ks, tstops)


def solve_gdre(prob, alg, **kw):
    """src/DifferentialRiccatiEquations.jl:78-94."""
    if isinstance(alg, Ros1):
        return solve_gdre_ros1(prob, alg, **kw)
    if isinstance(alg, Ros2):
        return solve_gdre_ros2(prob, alg, **kw)
    raise TypeError(alg)


def gare_residual(prob: GAREProblem, X: LDLt, *, AtL=None, EtL=None, BtLD=None, DLtGLD=None) -> LDLt:
    """src/riccati/residual.jl:6-52."""
    E, A, Q, G = prob.E, prob.A, prob.Q, prob.G
    if X.iszero():
        return copy.deepcopy(Q)
    gamma, Ct, S = Q.destructure()
    beta, B, Rinv = G.destructure()
    alpha, L, D = X.destructure()
    h, zk = Ct.shape[1], L.shape[1]
    dim = h + 2 * zk
    AtL = AtL if AtL is not None else A.T @ L
    EtL = EtL if EtL is not None else E.T @ L
    if DLtGLD is None:
        if BtLD is None:
            BtLD = (B.T @ L) @ D
            if alpha * beta != 1:
                BtLD = BtLD * (alpha * beta)
        DLtGLD = BtLD.T @ Rinv @ BtLD
    R = np.concatenate([Ct, AtL, EtL], axis=1)
    T = np.zeros((dim, dim))
    T[:h, :h] = gamma * S
    T[h:h + zk, h + zk:] = alpha * D
    T[h + zk:, h:h + zk] = T[h:h + zk, h + zk:]
    T[h + zk:, h + zk:] = -DLtGLD
    return compress(lowrank(R, T))


def gare_residual_dense(prob: GAREProblem, X: np.ndarray) -> np.ndarray:
    """src/riccati/residual.jl:54-66."""
    E, A, G, Q = prob.E, prob.A, prob.G, prob.Q
    alpha, B, D = G.destructure()
    Ed = E.toarray() if sp.issparse(E) else E
    Ad = A.toarray() if sp.issparse(A) else A
    BtXE = (B.T @ X) @ Ed
    return Q.to_dense() + Ad.T @ X @ Ed + Ed.T @ X @ Ad - BtXE.T @ (alpha * D) @ BtXE


def quadratic_forcing(_i, residual_norm):  # newton.jl:165
    return min(0.1, 0.9 * residual_norm)


def superlinear_forcing(i, _r):  # newton.jl:156
    return 1.0 / (i ** 3 + 1)


class Newton:  # src/riccati/types.jl:95-106
    def __init__(self, inner_alg=None, *, maxiters=5, reltol=None, abstol=None, inexact=True,
                 inexact_hybrid=True, inexact_forcing=quadratic_forcing, linesearch=True):
        self.inner_alg = inner_alg if inner_alg is not None else ADI()
        self.maxiters, self.reltol, self.abstol = maxiters, reltol, abstol
        self.inexact, self.inexact_hybrid = inexact, inexact_hybrid
        self.inexact_forcing, self.linesearch = inexact_forcing, linesearch


def solve_gare_newton(prob: GAREProblem, alg: Newton, *, observer=None) -> LDLt:
    """src/riccati/newton.jl:3-147."""
    _observe(observer, "observe_gare_start", prob, alg)
    E, A, Q = prob.E, prob.A, prob.Q
    alpha, B, _ = prob.G.destructure()
    assert alpha == 1
    alpha, Ct, _ = Q.destructure()
    assert alpha == 1
    res = Q
    res_norm = norm(res)
    reltol = alg.reltol if alg.reltol is not None else A.shape[0] * EPS
    abstol = alg.abstol if alg.abstol is not None else reltol * res_norm
    n = A.shape[1]
    X = lowrank(np.zeros((n, 0)), np.zeros((0, 0)))
    i = 0
    X_prev = None
    inner_alg = alg.inner_alg
    inner_reltol = inner_alg.reltol if getattr(inner_alg, "reltol", None) is not None else reltol / 10
    while True:
        alpha, L, D = X.destructure()
        EtL = E.T @ L
        BtLD = (B.T @ L) @ D
        if alpha != 1:
            BtLD = BtLD * alpha
        DLtGLD = BtLD.T @ BtLD
        K = BtLD @ EtL.T
        res = gare_residual(prob, X, EtL=EtL, DLtGLD=DLtGLD)
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
                    res = gare_residual(prob, X)
                    res_norm = norm(res)
                    if res_norm < (1 - lam * a_) * res_norm_prev:
                        alpha, L, D = X.destructure()
                        EtL = E.T @ L
                        BtLD = (B.T @ L) @ D
                        if alpha != 1:
                            BtLD = BtLD * alpha
                        DLtGLD = BtLD.T @ BtLD
                        K = BtLD @ EtL.T
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
        F = lr_update(A, -1.0, B, K)
        m = B.shape[1]
        q = Ct.shape[1]
        EtXB = EtL @ BtLD.T
        G = _hcat([Ct, EtXB])
        S = _dcat([np.eye(q), np.eye(m)])
        RHS = lowrank(G, S)
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
        X = solve_gale(lyap, inner_alg, abstol=inner_abstol, initial_guess=X_prev, observer=observer)
    _observe(observer, "observe_gare_done", i, X, res, res_norm)
    return X


# --------------------------------------------------------------------------------------------
# dense cross-checks (test references only: bartels-stewart.jl, dense_ros1.jl)
# --------------------------------------------------------------------------------------------
def bartels_stewart(prob: GALEProblem) -> np.ndarray:
    """src/lyapunov/bartels-stewart.jl:3-11 -- dense solution of A'XE + E'XA = -C."""
    E = prob.E.toarray() if sp.issparse(prob.E) else np.asarray(prob.E)
    A = prob.A.to_dense() if isinstance(prob.A, LowRankUpdate) else (
        prob.A.toarray() if sp.issparse(prob.A) else np.asarray(prob.A))
    C = prob.C.to_dense() if isinstance(prob.C, LDLt) else prob.C
    # A'XE + E'XA = -C  <=>  M'X + X M = -E^-T C E^-1  with M = A E^-1
    M = np.linalg.solve(E.T, A.T).T
    Q = np.linalg.solve(E.T, np.linalg.solve(E.T, C.T).T)
    X = sla.solve_continuous_lyapunov(M.T, -Q)
    return 0.5 * (X + X.T)


def dense_ros1(E, A, B, C, X0, tspan, dt):
    """src/riccati/dense_ros1.jl:3-55 -- dense implicit-Euler Rosenbrock reference (test/rail.jl:55)."""
    Ed = E.toarray() if sp.issparse(E) else E
    Ad = A.toarray() if sp.issparse(A) else A
    X = X0.copy()
    tstops = _tstops(tspan, dt)
    K = (B.T @ X) @ Ed
    Ks = [K]
    CtC = C.T @ C
    for i in range(1, len(tstops)):
        tau = tstops[i - 1] - tstops[i]
        F = (Ad - B @ K) - Ed / (2 * tau)
        R = CtC + K.T @ K + (1 / tau) * (Ed.T @ X @ Ed)
        X = bartels_stewart(GALEProblem(Ed, F, R))
        K = (B.T @ X) @ Ed
        Ks.append(K)
    return Ks, X
