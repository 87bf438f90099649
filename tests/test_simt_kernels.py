"""CPU tier: the product's CUDA kernel SOURCES (csrc/sparse_kernels.cu, csrc/dense_kernels.cu) executed on the
host-side SIMT emulator of tests/simt/ (every CUDA thread a fiber, warp collectives incl. the FP64 m8n8k4 MMA
emulated lane by lane, shared memory poisoned, cp.async deferred) with the product's launch schedule
(csrc/schedule.h) and symbolic analysis.  Catches indexing / fragment-layout / barrier mistakes without a GPU;
it says nothing about performance.  The GPU tier (tests/test_gpu_*.py) runs the same checks on the real thing."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import dre_b200
from tests import hostcheck
from tests.simt import emu


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


class _SymFromEmu(hostcheck.Sym):
    """hostcheck.Sym over the product's host-only symbolic analysis (same options as the emulated solver)."""

    def __init__(self, E, A, leaf, cap):
        from dre_b200 import capi

        super().__init__(capi.SymbolicAnalysis(E, A, leaf_size=(leaf or 96) | ((cap or 256) << 16)))


@pytest.mark.parametrize("n,leaf,cap,nrhs", [(371, 0, 0, 37), (1357, 24, 40, 13), (1357, 96, 256, 70)])
def test_emulated_factorization_and_sweeps(n, leaf, cap, nrhs):
    """k_assemble / k_extend_add / k_diag / k_l21 / k_schur / k_fwd / k_bwd, real and complex-symmetric:
    factor arrays entry by entry against the NumPy emulation of the algorithm, block solve against SuperLU."""
    E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
    S = emu.Solver(E, A, leaf, cap)
    H = _SymFromEmu(E, A, leaf, cap)
    assert H.info["nnz_L"] == S.sizes["nnz_L"] and H.nsn == S.sizes["nsn"]
    assert np.array_equal(H.perm, S.perm)
    rng = np.random.default_rng(5)
    a, e = 1.0, -1.0 / 200.0
    for mu in (-0.37, -0.02 + 0.11j):
        dtype = complex if isinstance(mu, complex) else float
        assert S.factor(a, e + mu) == 0
        Lh, Linvh, dvech = hostcheck.factor(H, a, e + mu, dtype)
        assert _rel(S.get("dvec"), dvech) < 1e-12
        Ld, Li = S.get("L"), S.get("Linv")
        for J in range(H.nsn):
            s_, f_ = H.s(J), H.s(J) + H.u(J)
            Pd = Ld[H.panel_off[J]:H.panel_off[J + 1]].reshape(s_, f_).T
            Ph = Lh[H.panel_off[J]:H.panel_off[J + 1]].reshape(s_, f_).T
            assert _rel(np.tril(Pd[:s_], -1), np.tril(Ph[:s_], -1)) < 1e-11 or np.linalg.norm(np.tril(Ph[:s_], -1)) == 0
            assert _rel(Pd[s_:], Ph[s_:]) < 1e-11 or np.linalg.norm(Ph[s_:]) == 0
            Lid = Li[H.linv_off[J]:H.linv_off[J + 1]].reshape(s_, s_).T
            assert _rel(np.tril(Lid), Linvh[J]) < 1e-11
        R = rng.standard_normal((n, nrhs))
        V = S.sweeps(R)
        M = (a * A + (e + mu) * E).tocsc().astype(dtype)
        assert _rel(V, spla.splu(M).solve(R.astype(dtype))) < 1e-11
        assert _rel(M @ V, R) < 1e-12
    S.close()


@pytest.mark.parametrize("n,leaf,cap,nrhs", [(371, 0, 0, 37), (1357, 24, 40, 150), (1357, 96, 256, 70),
                                             (5177, 24, 40, 37)])
def test_emulated_row_split_sweeps(n, leaf, cap, nrhs):
    """Sweep v2 (DRE_SWEEP2): k_m21 leaves M21 = L21 Linv in the panels, k_fwd2 / k_bwd2 spread the strips of a
    supernode over several CTAs.  (1357, 24, 40) has a populous leaf level (3 / 2 strips per warp) and sparse
    upper levels (1 strip per warp, 16- and 32-column chunks); (1357, 96, 256) has supernodes of several row blocks;
    (5177, 24, 40) has populous levels WITH children (11 levels, 494 supernodes: the 3-strip variant with the
    shared-memory tile of the children's contributions -- levels 1-4 of the n = 79 841 problem)."""
    E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
    S = emu.Solver(E, A, leaf, cap)
    H = _SymFromEmu(E, A, leaf, cap)
    rng = np.random.default_rng(8)
    a, e = 1.0, -1.0 / 200.0
    for mu in (-0.37, -0.02 + 0.11j):
        dtype = complex if isinstance(mu, complex) else float
        assert S.factor(a, e + mu, m21=True) == 0
        Lh, Linvh, dvech = hostcheck.factor(H, a, e + mu, dtype, m21=True)
        Ld = S.get("L")
        for J in range(H.nsn):
            s_, f_ = H.s(J), H.s(J) + H.u(J)
            Pd = Ld[H.panel_off[J]:H.panel_off[J + 1]].reshape(s_, f_).T
            Ph = Lh[H.panel_off[J]:H.panel_off[J + 1]].reshape(s_, f_).T
            assert _rel(Pd[s_:], Ph[s_:]) < 1e-11 or np.linalg.norm(Ph[s_:]) == 0
        R, Vt = rng.standard_normal((n, nrhs)), rng.standard_normal((n, 7))
        W = S.sweeps(R, Vt)
        M = (a * A + (e + mu) * E).tocsc().astype(dtype)
        RHS = np.hstack([R, Vt]).astype(dtype)
        assert _rel(W, spla.splu(M).solve(RHS)) < 1e-11
        assert _rel(M @ W, RHS) < 1e-12
    S.close()


def test_emulated_narrow_diag_kernel():
    """k_diag (first version, DRE_DIAG_V=1) with 64-thread CTAs (DRE_DIAG_NARROW_MIN): same factor arrays as its
    256-thread variant, bit for bit (the work distribution over warps changes, the arithmetic of every block does
    not); the default k_diag2 (block column in shared memory, inverse built inside the elimination loop) agrees with
    both to round-off, on chains of wide supernodes and on ragged last blocks."""
    n = 1357
    E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
    for leaf, cap in ((96, 256), (40, 72)):
        S = emu.Solver(E, A, leaf, cap)
        try:
            for mu in (-0.37, -0.02 + 0.11j):
                emu.set_diag_variant(1)
                emu.set_diag_narrow_min(1 << 30)
                assert S.factor(1.0, mu) == 0
                ref = [S.get(w).copy() for w in ("L", "Linv", "dvec")]
                emu.set_diag_narrow_min(1)
                assert S.factor(1.0, mu) == 0
                for w, r in zip(("L", "Linv", "dvec"), ref):
                    # (never-written entries above the diagonal blocks of Linv stay NaN-poisoned)
                    assert np.array_equal(S.get(w), r, equal_nan=True), w
                emu.set_diag_variant(2)
                assert S.factor(1.0, mu) == 0
                for w, r in zip(("L", "Linv", "dvec"), ref):
                    got = S.get(w)
                    assert np.array_equal(np.isnan(got), np.isnan(r)), w
                    ok = ~np.isnan(r)
                    assert np.max(np.abs(got[ok] - r[ok])) <= 1e-12 * np.max(np.abs(r[ok])), w
        finally:
            emu.set_diag_variant(2)
            emu.set_diag_narrow_min(1 << 30)
            S.close()


def test_emulated_sweeps_with_extra_rhs_panel():
    """The forward sweep reads [R, Vt] where the two panels lie (RhsSource): the SMW columns of the closed loop."""
    n = 371
    E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
    S = emu.Solver(E, A)
    rng = np.random.default_rng(6)
    R, Vt = rng.standard_normal((n, 21)), rng.standard_normal((n, 7))
    mu = -0.11
    assert S.factor(1.0, mu) == 0
    W = S.sweeps(R, Vt)
    M = (A + mu * E).tocsc()
    assert _rel(W, spla.splu(M).solve(np.hstack([R, Vt]))) < 1e-11
    S.close()


@pytest.mark.parametrize("variant", [1, 2])
def test_emulated_spmm(variant):
    """k_spmm and the opt-in k_spmm2 (DRE_SPMM2): 1 / 6 / 45 / 150 panel columns (one and two 128-column passes)."""
    n = 371
    E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
    S = emu.Solver(E, A)
    rng = np.random.default_rng(7)
    emu.set_spmm_variant(variant)
    try:
        for cols in (1, 6, 45, 150):
            X, Y = rng.standard_normal((n, cols)), rng.standard_normal((n, cols))
            assert _rel(S.spmm("E", -0.7, X, 1.0, Y), Y - 0.7 * (E @ X)) < 1e-13
            assert _rel(S.spmm("A", 2.0, X, 0.0, np.full_like(Y, np.nan)), 2.0 * (A @ X)) < 1e-13
    finally:
        emu.set_spmm_variant(1)
        S.close()


@pytest.mark.parametrize("n,a,b", [(500, 7, 70), (333, 12, 200), (700, 64, 64), (257, 45, 130), (130, 70, 300),
                                   (64, 3, 5)])
def test_emulated_gram(n, a, b):
    """k_gram_skinny (a <= 16), k_gram2<64> / k_gram2<128> (cp.async pipeline, DMMA) + k_reduce_partials."""
    rng = np.random.default_rng(a * 1000 + b)
    X, Y = rng.standard_normal((n, a)), rng.standard_normal((n, b))
    for sm in (148, 4):   # many / few row splits
        assert _rel(emu.gram(X, Y, sm_count=sm), X.T @ Y) < 1e-13
    w = rng.standard_normal(n)
    assert _rel(emu.gram(X, Y, roww=w), X.T @ (w[:, None] * Y)) < 1e-13


@pytest.mark.parametrize("n,a,b", [(300, 70, 64), (513, 130, 45), (200, 9, 64), (257, 40, 7), (129, 300, 130)])
def test_emulated_tall_gemm(n, a, b):
    """k_tall_gemm2<64> / <16> and the small-K variant, W plain and transposed, beta = 0 and accumulate."""
    rng = np.random.default_rng(a + b)
    X, W, Y = rng.standard_normal((n, a)), rng.standard_normal((a, b)), rng.standard_normal((n, b))
    assert _rel(emu.tall_gemm(1.5, X, W, 0.0, np.full_like(Y, np.nan)), 1.5 * X @ W) < 1e-13
    assert _rel(emu.tall_gemm(-1.0, X, W, 1.0, Y), Y - X @ W) < 1e-13
    assert _rel(emu.tall_gemm(1.0, X, np.ascontiguousarray(W.T), 1.0, Y, w_trans=True), Y + X @ W) < 1e-13


def test_emulated_small_dense_helpers():
    """k_pivchol (selection + inverse factor), k_norm_diag, k_colnorm2, k_cm2panel / k_panel2cm."""
    rng = np.random.default_rng(12)
    # pivoted Cholesky selection: a rank-5 Gram matrix of 40 columns; Q = P Wsel must be orthonormal and span P
    n, pb, rk = 300, 40, 5
    P = rng.standard_normal((n, rk)) @ rng.standard_normal((rk, pb))
    G = P.T @ P
    nsel, W, dfirst, remaining = emu.pivchol(G, 1e-20 * np.max(np.diag(G)), 1e-12)
    assert nsel == rk and abs(dfirst - np.max(np.diag(G))) < 1e-12 * dfirst
    Q = P @ W[:, :nsel]
    assert np.linalg.norm(Q.T @ Q - np.eye(nsel)) < 1e-8
    assert np.linalg.norm(P - Q @ (Q.T @ P)) < 1e-8 * np.linalg.norm(P)
    assert np.all(W[:, nsel:] == 0.0) and remaining < 1e-10 * dfirst
    # ||L diag(t) L'||_F^2 from G = L'L
    L = rng.standard_normal((n, 23))
    t = rng.standard_normal(23)
    ref = np.linalg.norm(L @ np.diag(t) @ L.T) ** 2
    assert abs(emu.norm_diag(L.T @ L, t) - ref) < 1e-12 * ref
    # column norms
    M = rng.standard_normal((n, 37))
    assert np.allclose(emu.colnorm2(M), np.sum(M * M, axis=0), rtol=1e-13)
    # upload / download transposes with the row permutation
    iperm = rng.permutation(n).astype(np.int32)
    panel, back = emu.panel_roundtrip(M, iperm)
    assert np.array_equal(panel[iperm], M) and np.array_equal(back, M)


def test_emulated_row_split_sweeps_fat_populous_levels(monkeypatch):
    """Every level treated as populous (DRE_SWEEP2_POPULOUS_MIN=1) on wide supernodes: the 3-strip forward / 2-strip
    backward variants with fronts of more than 192 rows, i.e. several row blocks per supernode and a non-zero offset
    into the children's contribution tile."""
    monkeypatch.setenv("DRE_SWEEP2_POPULOUS_MIN", "1")
    n = 20209                      # 9 levels, supernodes up to 141 columns, fronts of ~300 rows
    E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
    S = emu.Solver(E, A, 96, 256)
    rng = np.random.default_rng(9)
    R = rng.standard_normal((n, 8))
    for mu in (-0.37, -0.02 + 0.11j):
        dtype = complex if isinstance(mu, complex) else float
        assert S.factor(1.0, mu, m21=True) == 0
        W = S.sweeps(R)
        M = (A + mu * E).tocsc().astype(dtype)
        assert _rel(M @ W, R.astype(dtype)) < 1e-12
    S.close()
