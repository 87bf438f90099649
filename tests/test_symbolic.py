"""Host-side symbolic analysis (C++, through the C ABI, no GPU): the exported structures are
validated by a NumPy emulation of the device algorithms (tests/hostcheck.py) against SciPy."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import dre_b200
from dre_b200 import capi
from tests import hostcheck

pencils = dre_b200.pencils


def sa_export(S, name):
    return S._sa.export(name)


def _check_structure(S):
    n = S.n
    assert sorted(S.perm.tolist()) == list(range(n))
    assert np.array_equal(S.iperm[S.perm], np.arange(n))
    assert S.sn_first[0] == 0 and S.sn_first[-1] == n and np.all(np.diff(S.sn_first) > 0)
    for J in range(S.nsn):
        rows = S.sn_rows[S.sn_rowptr[J]:S.sn_rowptr[J + 1]]
        assert np.all(np.diff(rows) > 0)
        if len(rows):
            assert rows[0] >= S.sn_first[J + 1]
            P = S.sn_parent[J]
            assert P > J and S.sn_first[P] <= rows[0] < S.sn_first[P + 1]
            assert S.sn_level[P] == S.sn_level[J] + 1
            rel = S.relmap[S.sn_rowptr[J]:S.sn_rowptr[J + 1]]
            assert np.all(np.diff(rel) > 0) and rel[-1] < S.s(P) + S.u(P)
        else:
            assert S.sn_parent[J] == -1
    # device processing order: bottom subtrees (ascending lists), then top levels; every child must be
    # finished before its parent, and a subtree must contain all descendants of its root
    sub = sa_export(S, "sn_subtree")
    st_ptr, st_sn = sa_export(S, "st_ptr"), sa_export(S, "st_sn")
    tl_ptr, tl_sn = sa_export(S, "top_level_ptr"), sa_export(S, "top_level_sn")
    order = {}
    k = 0
    for t in range(len(st_ptr) - 1):
        lst = st_sn[st_ptr[t]:st_ptr[t + 1]]
        assert np.all(np.diff(lst) > 0) and np.all(sub[lst] == t)
        for J in lst:
            order[int(J)] = (0, t, k)
            k += 1
    for l in range(len(tl_ptr) - 1):
        for J in tl_sn[tl_ptr[l]:tl_ptr[l + 1]]:
            assert sub[J] == -1
            order[int(J)] = (1, l, 0)
    assert len(order) == S.nsn
    for J in range(S.nsn):
        P = S.sn_parent[J]
        if P < 0:
            continue
        if sub[P] >= 0:
            assert sub[J] == sub[P] and order[J][2] < order[P][2]
        elif sub[J] >= 0:
            pass  # subtree root below a top supernode: bottom phase runs first
        else:
            assert order[J][1] < order[P][1]
    assert S.s(0) <= 32 and max(S.s(J) for J in range(S.nsn)) <= 32


@pytest.mark.parametrize("n", [371, 1357])
def test_rail_factor_solve_real_and_complex(n):
    E, A, B, C, _ = pencils.rail_pencil(n)
    sa = capi.SymbolicAnalysis(E, A, leaf_size=32)
    S = hostcheck.Sym(sa)
    _check_structure(S)
    rng = np.random.default_rng(0)
    Rhs = rng.standard_normal((n, 5))
    a, e = 1.0, -1.0 / 200.0
    for mu in (-0.37, -0.02 + 0.11j):
        dtype = complex if isinstance(mu, complex) else float
        L, dblk = hostcheck.factor(S, a, e + mu, dtype)
        X = hostcheck.solve(S, L, dblk, Rhs[S.perm])
        M = (a * A + (e + mu) * E).tocsc()
        Xref = spla.splu(M.astype(dtype)).solve(Rhs.astype(dtype))
        err = np.linalg.norm(X[S.iperm] - Xref) / np.linalg.norm(Xref)
        assert err < 1e-10, err
    # permuted CSR copies used by the SpMM kernels
    Ep = hostcheck.permuted_csr(S, "E")
    assert abs(Ep - E[S.perm][:, S.perm]).max() < 1e-300 + 1e-16 * abs(E).max()
    Ap = hostcheck.permuted_csr(S, "A")
    assert abs(Ap - A[S.perm][:, S.perm]).max() < 1e-300 + 1e-16 * abs(A).max()


def test_random_and_disconnected_pencils():
    E, A = pencils.random_spd_pencil(50, seed=3)
    S = hostcheck.Sym(capi.SymbolicAnalysis(E, A, leaf_size=8))
    _check_structure(S)
    L, dblk = hostcheck.factor(S, 1.0, -0.7, float)
    b = np.arange(50.0).reshape(50, 1)
    X = hostcheck.solve(S, L, dblk, b[S.perm])
    ref = spla.spsolve((A - 0.7 * E).tocsc(), b)
    assert np.allclose(X[S.iperm, 0], ref, rtol=1e-10)
    # diagonal (fully disconnected) pencil
    n = 100
    E = sp.identity(n, format="csc") * 2.0
    A = -sp.diags(np.arange(1.0, n + 1)).tocsc()
    S = hostcheck.Sym(capi.SymbolicAnalysis(E, A, leaf_size=16))
    _check_structure(S)
    L, dblk = hostcheck.factor(S, 1.0, -1.0, float)
    X = hostcheck.solve(S, L, dblk, np.ones((n, 1)))
    assert np.allclose(X[S.iperm, 0], 1.0 / (-np.arange(1.0, n + 1) - 2.0))


def test_heat3d_structure():
    E, A, B, C, _ = pencils.heat3d_pencil(12)
    sa = capi.SymbolicAnalysis(E, A, leaf_size=32)
    S = hostcheck.Sym(sa)
    _check_structure(S)
    L, dblk = hostcheck.factor(S, 1.0, -3.0, float)
    b = np.ones((S.n, 2))
    X = hostcheck.solve(S, L, dblk, b)
    ref = spla.splu((A - 3.0 * E).tocsc()).solve(b)
    assert np.linalg.norm(X[S.iperm] - ref) / np.linalg.norm(ref) < 1e-10


def test_nonsymmetric_rejected():
    E, A = pencils.random_spd_pencil(30, seed=1)
    A = A.tolil()
    A[0, 5] = 0.3
    with pytest.raises(capi.DreError, match="symmetric"):
        capi.SymbolicAnalysis(E, A.tocsc())


def test_library_exports_every_declared_symbol():
    import re, os
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(__file__)), "include", "dre_b200.h")).read()
    declared = set(re.findall(r"DRE_API [\w\s\*]+?\b(dre_\w+)\(", hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS)
    lib = capi.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.dre_version()
