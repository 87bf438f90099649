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
            assert S.sn_level[P] > S.sn_level[J]
            rel = S.relmap[S.sn_rowptr[J]:S.sn_rowptr[J + 1]]
            assert np.all(np.diff(rel) > 0) and rel[-1] < S.s(P) + S.u(P)
        else:
            assert S.sn_parent[J] == -1
    # level schedule: every child sits on a strictly lower level than its parent; leaves are level 0
    for J in range(S.nsn):
        P = S.sn_parent[J]
        if P >= 0:
            assert S.sn_level[J] < S.sn_level[P]
        if S.child_ptr[J + 1] == S.child_ptr[J]:
            assert S.sn_level[J] == 0
    assert sorted(S.level_sn.tolist()) == list(range(S.nsn))
    assert max(S.s(J) for J in range(S.nsn)) <= 256
    assert S.linv_off[-1] == sum(S.s(J) ** 2 for J in range(S.nsn))
    # update matrices / update vectors are pooled by liveness: supernode J owns its region from level(J) to
    # level(parent(J)); regions of supernodes that are alive at the same time never overlap, the pool is no larger
    # than the layout without reuse
    assert S.upd_off[-1] <= sum(S.u(J) ** 2 for J in range(S.nsn))
    for off, size, total in ((S.upd_off, lambda J: S.u(J) ** 2, int(S.upd_off[-1])),
                             (S.rhs_off, lambda J: S.u(J), None)):
        iv = []
        for J in range(S.nsn):
            P = S.sn_parent[J]
            if size(J) == 0:
                continue
            iv.append((int(S.sn_level[J]), int(S.sn_level[P]) if P >= 0 else int(S.sn_level[J]), int(off[J]),
                       int(off[J]) + size(J)))
            assert total is None or iv[-1][3] <= total
        for l in range(S.nlevels):
            alive = sorted((a, b) for (l0, l1, a, b) in iv if l0 <= l <= l1)
            for (a0, b0), (a1, b1) in zip(alive, alive[1:]):
                assert b0 <= a1, "live regions overlap"


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
        L, Linv, dvec = hostcheck.factor(S, a, e + mu, dtype)
        X = hostcheck.solve(S, L, Linv, dvec, Rhs[S.perm])
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
    L, Linv, dvec = hostcheck.factor(S, 1.0, -0.7, float)
    b = np.arange(50.0).reshape(50, 1)
    X = hostcheck.solve(S, L, Linv, dvec, b[S.perm])
    ref = spla.spsolve((A - 0.7 * E).tocsc(), b)
    assert np.allclose(X[S.iperm, 0], ref, rtol=1e-10)
    # diagonal (fully disconnected) pencil
    n = 100
    E = sp.identity(n, format="csc") * 2.0
    A = -sp.diags(np.arange(1.0, n + 1)).tocsc()
    S = hostcheck.Sym(capi.SymbolicAnalysis(E, A, leaf_size=16))
    _check_structure(S)
    L, Linv, dvec = hostcheck.factor(S, 1.0, -1.0, float)
    X = hostcheck.solve(S, L, Linv, dvec, np.ones((n, 1)))
    assert np.allclose(X[S.iperm, 0], 1.0 / (-np.arange(1.0, n + 1) - 2.0))


def test_heat3d_structure():
    E, A, B, C, _ = pencils.heat3d_pencil(12)
    sa = capi.SymbolicAnalysis(E, A, leaf_size=32)
    S = hostcheck.Sym(sa)
    _check_structure(S)
    L, Linv, dvec = hostcheck.factor(S, 1.0, -3.0, float)
    b = np.ones((S.n, 2))
    X = hostcheck.solve(S, L, Linv, dvec, b)
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


def test_threaded_dissection_is_deterministic(monkeypatch):
    """The nested dissection runs the two halves of its top levels as parallel tasks (DRE_SYMBOLIC_THREADS); the
    ordering, the supernode partition and every index map must not depend on the number of threads."""
    E, A, B, C, _ = pencils.rail_pencil(20209)
    names = ("perm", "sn_first", "sn_rows", "relmap", "level_sn", "asm_dest", "upd_off", "rhs_off")
    ref = None
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("DRE_SYMBOLIC_THREADS", threads)
        sa = capi.SymbolicAnalysis(E, A)
        cur = {k: sa.export(k) for k in names}
        if ref is None:
            ref = cur
        else:
            for k in names:
                assert np.array_equal(ref[k], cur[k]), (threads, k)


def test_pencil_key_is_content_based():
    """the resident pencil is recognised by content (an equal copy does not trigger a re-upload that would
    invalidate every DeviceMatrix; values changed in place are noticed)."""
    from dre_b200 import api

    E, A, _, _, _ = dre_b200.pencils.rail_pencil(371)
    k = api.Backend._pencil_key(E, A)
    assert api.Backend._pencil_key(E.copy(), A.copy()) == k
    E2 = E.copy()
    E2.data[3] *= 1.0 + 1e-12
    assert api.Backend._pencil_key(E2, A) != k
    A2 = A.copy()
    A2.indices[[0, 1]] = A2.indices[[1, 0]]
    assert api.Backend._pencil_key(E, A2) != k
