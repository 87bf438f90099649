"""GPU unit tests of the C-ABI primitives (run on the B200 box: pytest -m gpu).  Each primitive is
compared with NumPy/SciPy on the same seeded inputs; tolerances are written next to each check."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import dre_b200
from dre_b200 import api

pytestmark = pytest.mark.gpu
pencils = dre_b200.pencils


_RAIL = pencils.rail_pencil(1357)[:4]


@pytest.fixture()
def rail():
    E, A, B, C = _RAIL
    api.upload_pencil(E, A)  # no-op while the same pencil is resident
    return E, A, B, C


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_upload_download_roundtrip(rail):
    E, A, B, C = rail
    rng = np.random.default_rng(0)
    for k in (1, 7, 33, 130):
        M = rng.standard_normal((E.shape[0], k))
        d = api.DeviceMatrix.from_host(M)
        assert np.array_equal(d.to_host(), M)  # bit exact
        assert np.array_equal(d.cols(k // 2, k).to_host(), M[:, k // 2:])
        assert np.array_equal(d.copy().to_host(), M)


def test_spmm_matches_scipy(rail):
    E, A, B, C = rail
    rng = np.random.default_rng(1)
    for k in (1, 5, 64, 157):
        X = rng.standard_normal((E.shape[0], k))
        Y0 = rng.standard_normal((E.shape[0], k))
        Xd, Yd = api.DeviceMatrix.from_host(X), api.DeviceMatrix.from_host(Y0)
        assert _rel(api.spmm("E", Xd).to_host(), E @ X) < 1e-14
        assert _rel(api.spmm("A", Xd).to_host(), A @ X) < 1e-14
        api.spmm("E", Xd, -0.7, Yd, 1.0)
        assert _rel(Yd.to_host(), Y0 - 0.7 * (E @ X)) < 1e-14


def test_gram_and_tall_gemm(rail):
    E, A, B, C = rail
    n = E.shape[0]
    rng = np.random.default_rng(2)
    for a, b in ((1, 1), (7, 164), (65, 130), (157, 157), (200, 3)):
        X = rng.standard_normal((n, a))
        Y = rng.standard_normal((n, b))
        Xd, Yd = api.DeviceMatrix.from_host(X), api.DeviceMatrix.from_host(Y)
        G = api.gemm_tn(Xd, Yd)
        assert _rel(G, X.T @ Y) < 1e-13, (a, b)
        W = rng.standard_normal((a, b))
        Z = api.gemm_nn(Xd, W, 0.5, Yd, -2.0)
        assert _rel(Z.to_host(), 0.5 * X @ W - 2.0 * Y) < 1e-13, (a, b)
        Z2 = api.gemm_nn(Xd, W)
        assert _rel(Z2.to_host(), X @ W) < 1e-13
    # views into a wider panel (odd offsets)
    P = rng.standard_normal((n, 50))
    Pd = api.DeviceMatrix.from_host(P)
    G = api.gemm_tn(Pd.cols(3, 20), Pd.cols(21, 50))
    assert _rel(G, P[:, 3:20].T @ P[:, 21:50]) < 1e-13


@pytest.mark.parametrize("mu", [-0.37, -3.1e-3, -0.02 + 0.11j, -4.0 - 0.5j])
def test_shift_solve_plain(rail, mu):
    """(a A + e E + mu E) V = R against SciPy SuperLU; tolerance 1e-10 relative (kappa ~ 1e5)."""
    E, A, B, C = rail
    n = E.shape[0]
    rng = np.random.default_rng(3)
    R = rng.standard_normal((n, 37))
    a, e = 1.0, -1.0 / 200.0
    F = api.PencilCombo(a, e)
    out = api.solve_block(api.BlockLinearProblem(F, api.DeviceMatrix.from_host(R)), mu=mu)
    M = (a * A + (e + mu) * E).tocsc()
    ref = spla.splu(M.astype(complex if isinstance(mu, complex) else float)).solve(
        R.astype(complex if isinstance(mu, complex) else float))
    if isinstance(mu, complex):
        V = out[0].to_host() + 1j * out[1].to_host()
    else:
        V = out.to_host()
    assert _rel(V, ref) < 1e-10
    assert _rel(M @ V, R) < 1e-11


@pytest.mark.parametrize("n,leaf,cap", [(1357, 96, 256), (5177, 96, 256), (5177, 24, 40), (20209, 128, 256)])
def test_factor_arrays_and_wide_solves(n, leaf, cap, monkeypatch):
    """The numeric factorization on the device (panels L, explicit inverse diagonal blocks, pivots) against
    the NumPy emulation of the same algorithm on the same symbolic structure (tests/hostcheck.py), entry by
    entry, and the block solve with 250 / 37 / 3 right-hand sides against SuperLU; real and complex shifts.
    (leaf, cap) = (24, 40) forces chains of narrow supernodes, (128, 256) the widest leaves."""
    from dre_b200 import capi
    from tests import hostcheck

    monkeypatch.setenv("DRE_LEAF_SIZE", str(leaf))
    monkeypatch.setenv("DRE_MAX_SNODE", str(cap))
    E, A, B, C, _ = pencils.rail_pencil(n)
    api.reset_backend()
    api.upload_pencil(E, A)
    be = api.backend()
    S = hostcheck.Sym(capi.SymbolicAnalysis(E, A, leaf_size=leaf | (cap << 16)))
    info = be.ctx.symbolic_info()
    assert info["nnz_L"] == S.info["nnz_L"] and info["nsupernodes"] == S.nsn
    rng = np.random.default_rng(11)
    a, e = 1.0, -1.0 / 200.0
    F = api.PencilCombo(a, e)
    for mu in (-0.37, -0.02 + 0.11j):
        cx = isinstance(mu, complex)
        dtype = complex if cx else float
        for nrhs in (250, 37, 3):
            R = rng.standard_normal((n, nrhs))
            out = api.solve_block(api.BlockLinearProblem(F, api.DeviceMatrix.from_host(R)), mu=mu)
            V = out[0].to_host() + 1j * out[1].to_host() if cx else out.to_host()
            if nrhs == 250:
                Lh, Linvh, dvech = hostcheck.factor(S, a, e + mu, dtype)
                dv = be.ctx.debug_export("dvec", cx)
                assert _rel(dv, dvech) < 1e-11, "pivots"
                Ld = be.ctx.debug_export("L", cx)
                for J in range(S.nsn):
                    s_, f_ = S.s(J), S.s(J) + S.u(J)
                    Pd = Ld[S.panel_off[J]:S.panel_off[J + 1]].reshape(s_, f_).T
                    Ph = Lh[S.panel_off[J]:S.panel_off[J + 1]].reshape(s_, f_).T
                    assert _rel(np.tril(Pd[:s_], -1), np.tril(Ph[:s_], -1)) < 1e-10 or \
                        np.linalg.norm(np.tril(Ph[:s_], -1)) < 1e-300, ("L11", J)
                    assert _rel(Pd[s_:], Ph[s_:]) < 1e-10 or np.linalg.norm(Ph[s_:]) < 1e-300, ("L21", J)
                Li = be.ctx.debug_export("Linv", cx)
                for J in range(S.nsn):
                    s_ = S.s(J)
                    Lid = Li[S.linv_off[J]:S.linv_off[J + 1]].reshape(s_, s_).T
                    assert _rel(np.tril(Lid), Linvh[J]) < 1e-10, ("Linv", J)
            M = (a * A + (e + mu) * E).tocsc()
            ref = spla.splu(M.astype(dtype)).solve(R.astype(dtype))
            assert _rel(V, ref) < 1e-10, (mu, nrhs)
            assert _rel(M @ V, R) < 1e-11, (mu, nrhs)
    api.reset_backend()


def test_full_size_solve_and_compress_properties():
    """BASELINE's full size (n = 79 841, 250 right-hand sides), checked through size-independent properties
    because SuperLU / dense references are out of reach here:  the residual of the shifted block solve
    ||(a A + (e+mu) E) V - R|| / ||R|| computed with the device SpMM (real and complex shift), and the
    compress! invariants -- value preserved under a second compression, orthonormal outer factor, norm
    unchanged (test/LDLt.jl:63-90)."""
    n = 79841
    E, A, B, C, _ = pencils.rail_pencil(n)
    api.reset_backend()
    api.upload_pencil(E, A)
    rng = np.random.default_rng(21)
    R = rng.standard_normal((n, 250))
    Rd = api.DeviceMatrix.from_host(R)
    a, e = 1.0, -1.0 / 200.0
    F = api.PencilCombo(a, e)
    for mu in (-0.37, -0.02 + 0.11j):
        out = api.solve_block(api.BlockLinearProblem(F, Rd), mu=mu)
        parts = out if isinstance(mu, complex) else (out,)
        res = []
        for part, (cr, ci) in zip(parts, ((1.0, 0.0), (0.0, 1.0))):
            res.append((part, cr, ci))
        # real part of M V:  a A Vr + Re(e+mu) E Vr - Im(mu) E Vi ;  imaginary part analogously
        Vr = parts[0]
        Vi = parts[1] if isinstance(mu, complex) else None
        mr, mi = (e + mu.real, mu.imag) if isinstance(mu, complex) else (e + mu, 0.0)
        Yr = api.spmm("A", Vr, a)
        api.spmm("E", Vr, mr, Yr, 1.0)
        if Vi is not None:
            api.spmm("E", Vi, -mi, Yr, 1.0)
        resid = Yr.to_host() - R
        assert np.linalg.norm(resid) <= 1e-10 * np.linalg.norm(R), mu
        if Vi is not None:
            Yi = api.spmm("A", Vi, a)
            api.spmm("E", Vi, mr, Yi, 1.0)
            api.spmm("E", Vr, mi, Yi, 1.0)
            assert np.linalg.norm(Yi.to_host()) <= 1e-10 * np.linalg.norm(R), mu
    # compress! invariants on a rank-deficient 250-column factor with an indefinite diagonal core
    base = rng.standard_normal((n, 40))
    L = base @ rng.standard_normal((40, 120))
    d = rng.standard_normal(120)
    X = api.lowrank(api.DeviceMatrix.from_host(L), np.diag(d))
    nrm0 = api.norm(X)
    api.compress_(X)
    k1 = X.rank()
    assert 38 <= k1 <= 42
    assert abs(api.norm(X) - nrm0) <= 1e-10 * nrm0
    G = api.gemm_tn(X.Ls[0], X.Ls[0])
    assert np.linalg.norm(G - np.eye(k1)) < 1e-10
    lam1 = np.sort(np.diag(X.Ds[0]))
    api.compress_(X)  # idempotent up to round-off (uses the orthonormal hint of the first result)
    assert X.rank() == k1
    assert np.allclose(np.sort(np.diag(X.Ds[0])), lam1, rtol=1e-9, atol=1e-9 * np.abs(lam1).max())
    api.reset_backend()


@pytest.mark.parametrize("mu", [-0.2, -0.05 + 0.3j])
def test_shift_solve_lowrank_smw(rail, mu):
    """Closed-loop operator (A_s - B K + mu E) with the fused SMW correction
    (reference: test/LowRankUpdate.jl:31-40, M*X ~ B)."""
    E, A, B, C = rail
    n, m = B.shape
    rng = np.random.default_rng(4)
    K = rng.standard_normal((m, n)) * 1e-3
    R = rng.standard_normal((n, 29))
    a, e = 1.0, -0.005
    F = api.LowRankUpdate(api.PencilCombo(a, e), -1.0, api.DeviceMatrix.from_host(B), api.DeviceMatrix.from_host(K.T))
    out = api.solve_block(api.BlockLinearProblem(F, api.DeviceMatrix.from_host(R)), mu=mu)
    V = out[0].to_host() + 1j * out[1].to_host() if isinstance(mu, complex) else out.to_host()
    Md = (a * A + (e + mu) * E).toarray() - B @ K
    assert _rel(Md @ V, R) < 1e-10
    # transposed operator as used by ADI: (F' + mu E') V = R
    api._set_operator(F, transpose=True)
    be = api.backend()
    Rd = api.DeviceMatrix.from_host(R)
    V1, V2 = api.DeviceMatrix.empty(29), api.DeviceMatrix.empty(29)
    mu_c = complex(mu)
    be.check(be.lib.dre_shift_solve(be.h, mu_c.real, mu_c.imag, Rd.view, V1.view, V2.view))
    Vt = V1.to_host() + (1j * V2.to_host() if mu_c.imag else 0)
    assert _rel(Md.T @ Vt, R) < 1e-10


def test_norm_diag_and_dense(rail):
    """norm(::LDLt) (src/LDLt.jl:77-89); reference check: norm(X) ~ norm(Matrix(X)) (test/LDLt.jl:63)."""
    E, A, B, C = rail
    n = E.shape[0]
    rng = np.random.default_rng(5)
    L = rng.standard_normal((n, 40)) * np.logspace(0, -6, 40)
    d = rng.standard_normal(40)
    X = api.lowrank(api.DeviceMatrix.from_host(L), np.diag(d))
    ref = np.linalg.norm(L @ np.diag(d) @ L.T)
    assert abs(api.norm(X) - ref) < 1e-12 * ref
    assert abs(api.norm(-2.5 * X) - 2.5 * ref) < 1e-12 * ref
    S = rng.standard_normal((40, 40))
    S = S + S.T
    X = api.lowrank(api.DeviceMatrix.from_host(L), S)
    ref = np.linalg.norm(L @ S @ L.T)
    assert abs(api.norm(X) - ref) < 1e-12 * ref
    Y = X + X
    assert abs(api.norm(Y) - 2 * ref) < 1e-12 * ref


def test_compress_matches_dense(rail):
    """compress! (src/LDLt.jl:204-225): value preserved, rank revealed, orthonormal factor, diagonal core
    (test/LDLt.jl:76-90)."""
    E, A, B, C = rail
    n = E.shape[0]
    rng = np.random.default_rng(6)
    base = rng.standard_normal((n, 30))
    L1 = base @ rng.standard_normal((30, 70))            # rank 30 stored in 70 columns
    L2 = base[:, :10] @ rng.standard_normal((10, 25)) + 1e-9 * rng.standard_normal((n, 25))
    D1 = np.diag(rng.standard_normal(70))
    S2 = rng.standard_normal((25, 25))
    S2 = S2 + S2.T
    X = 1.5 * api.lowrank(api.DeviceMatrix.from_host(L1), D1) + (-0.5) * api.lowrank(api.DeviceMatrix.from_host(L2), S2)
    dense = 1.5 * L1 @ D1 @ L1.T - 0.5 * L2 @ S2 @ L2.T
    api.compress_(X)
    assert len(X.Ls) == 1 and X.alphas == [1.0]
    Lc, Dc = X.Ls[0].to_host(), X.Ds[0]
    assert np.count_nonzero(Dc - np.diag(np.diag(Dc))) == 0
    assert _rel(Lc @ Dc @ Lc.T, dense) < 1e-12
    assert np.linalg.norm(Lc.T @ Lc - np.eye(Lc.shape[1])) < 1e-11
    lam = np.linalg.eigvalsh(dense)
    expect = np.sum(np.abs(lam) >= 100 * np.max(np.abs(lam)) * np.finfo(float).eps)
    assert abs(X.rank() - expect) <= 2  # eigenvalues at the truncation threshold may fall either side
    # rank-1 core (test/LDLt.jl:84-89)
    S = np.zeros((25, 25))
    S[0, 0] = 13.0
    Y = api.compress_(api.lowrank(api.DeviceMatrix.from_host(L2), S))
    assert Y.rank() == 1

def test_rrqr_orth(rail):
    E, A, B, C = rail
    n = E.shape[0]
    rng = np.random.default_rng(7)
    N1 = rng.standard_normal((n, 20))
    N2 = np.concatenate([N1[:, :5] @ rng.standard_normal((5, 15)), rng.standard_normal((n, 3)) * 1e-7], axis=1)
    Et, At = api.orth_restrict([api.DeviceMatrix.from_host(N1), api.DeviceMatrix.from_host(N2)],
                               api.PencilCombo(0, 1), api.PencilCombo(1, 0))
    import scipy.linalg as sla
    N = np.concatenate([N1, N2], axis=1)
    U, s, _ = sla.svd(N, full_matrices=False)
    Q = U[:, s > n * np.finfo(float).eps]
    assert Et.shape[0] == Q.shape[1] == 23
    lam = np.sort_complex(sla.eigvals(At, Et))
    ref = np.sort_complex(sla.eigvals(Q.T @ (A @ Q), Q.T @ (E @ Q)))
    assert np.allclose(lam, ref, rtol=1e-8)


@pytest.mark.parametrize("k", [1, 2, 3, 5, 31, 64, 130, 300, 517])
def test_symmetric_eigensolver_vs_lapack(k):
    """The in-tree eigensolver behind compress! (src/LDLt.jl:214 calls LAPACK): backward error and orthogonality
    at LAPACK level on the kinds of cores this path produces -- indefinite with +-pairs (T = [0 D; D 0] of the
    Lyapunov residual, src/lyapunov/residual.jl:21-28), eigenvalues decaying over 16 decades, exact rank
    deficiency, repeated eigenvalues."""
    ctx = api.backend().ctx
    rng = np.random.default_rng(k)
    cases = []
    Q, _ = np.linalg.qr(rng.standard_normal((k, k)))
    cases.append((Q * np.logspace(0, -16, k) * rng.choice([-1.0, 1.0], k)) @ Q.T)          # graded, indefinite
    h = k // 2
    if h:
        D = np.diag(rng.standard_normal(h))
        Tm = np.zeros((k, k))
        Tm[:h, h:2 * h] = D
        Tm[h:2 * h, :h] = D
        cases.append(Q @ Tm @ Q.T)                                                          # exact +- pairs (+ a zero)
        cases.append(Tm)                                                                    # already sparse / reducible
    G = rng.standard_normal((k, max(1, k // 3)))
    cases.append(G @ G.T)                                                                   # rank deficient
    cases.append(np.eye(k) * 3.0 + 1e-9 * (Q + Q.T))                                        # clustered
    for S in cases:
        S = 0.5 * (S + S.T)
        w, V = ctx.debug_eigh(S)
        nrm = max(np.linalg.norm(S, 2), 1e-300)
        assert np.all(np.diff(w) >= 0)
        assert np.linalg.norm(V.T @ V - np.eye(k)) <= 50 * k * 2.2e-16
        assert np.linalg.norm(S @ V - V * w) <= 50 * k * 2.2e-16 * nrm
        assert np.max(np.abs(w - np.linalg.eigvalsh(S))) <= 50 * k * 2.2e-16 * nrm
