"""CPU tier (SIMT emulator): the reference's own unit tests for the types of the hot path, run through the product API
(api.* on top of the emulated C ABI -- the same CUDA sources compiled for the host).  Each test names the reference
test it mirrors: test/LDLt.jl, test/residual.jl, test/LowRankUpdate.jl, test/runtests.jl:12-19.  Random inputs are
seeded here (the reference's are not)."""
import numpy as np
import pytest

import dre_b200
from dre_b200 import api, capi
from tests.simt import build_emu


@pytest.fixture()
def emulated(monkeypatch):
    saved = (capi.LIB_PATH, capi._lib)
    monkeypatch.setenv("DRE_NO_PRIME", "1")
    api.reset_backend()
    capi.LIB_PATH, capi._lib = build_emu.build(), None
    api.reset_backend()
    n = 371
    E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
    api.upload_pencil(E, A)
    yield n, E, A, B, C
    api.reset_backend()
    capi.LIB_PATH, capi._lib = saved


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_ldlt_type_semantics(emulated):
    """test/LDLt.jl:21-74 -- size / rank, Matrix conversion, destructuring returns the SAME objects, scalar and additive
    arithmetic, norm(2X) = 2 norm(X), norm(X) = norm(Matrix(X)), X + zero == X."""
    n, *_ = emulated
    rng = np.random.default_rng(0)
    k = 2
    L = rng.standard_normal((n, k))
    D = rng.standard_normal((k, k))
    D = D + D.T
    Ld = api.DeviceMatrix.from_host(L)
    X = api.lowrank(Ld, D)
    assert X.shape == (n, n) and X.rank() == k
    M = L @ D @ L.T
    assert _rel(X.to_dense(), M) < 1e-14                          # :21-27
    a, L_, D_ = X.destructure()
    assert a == 1.0 and L_ is Ld and np.array_equal(D_, D)        # :29-35 (same outer factor object)
    assert _rel((2 * X).to_dense(), 2 * M) < 1e-14
    assert _rel((X / 4).to_dense(), M / 4) < 1e-14
    assert _rel((-X).to_dense(), -M) < 1e-14
    assert _rel((X + X).to_dense(), 2 * M) < 1e-14
    assert _rel((X - 0.25 * X).to_dense(), 0.75 * M) < 1e-14
    assert abs(api.norm(2 * X) - 2 * api.norm(X)) <= 1e-13 * api.norm(X)        # :55-58
    assert abs(api.norm(X) - np.linalg.norm(M)) <= 1e-12 * np.linalg.norm(M)    # :59-60
    Z = X.zero()
    assert Z.iszero() and Z.rank() == 0 and api.norm(Z) == 0.0
    assert (X + Z) is X and (Z + X) is X                                          # :62-74
    # identity core by default (lowrank(L) == L L')
    assert _rel(api.lowrank(Ld).to_dense(), L @ L.T) < 1e-14


def test_compress_keeps_rank_and_value(emulated):
    """test/LDLt.jl:76-90 -- compress!(X + X) has rank k and the value 2X; a rank-1 core compresses to rank 1."""
    n, *_ = emulated
    rng = np.random.default_rng(1)
    k = 2
    L = rng.standard_normal((n, k))
    D = np.diag([1.5, -0.5])
    X = api.lowrank(api.DeviceMatrix.from_host(L), D)
    Y = api.compress_(X + X)
    assert Y.rank() == k
    assert _rel(Y.to_dense(), 2 * L @ D @ L.T) < 1e-13
    L3 = rng.standard_normal((n, 3))
    S = np.zeros((3, 3))
    S[1, 1] = 7.0
    W = api.compress_(api.lowrank(api.DeviceMatrix.from_host(L3), S))
    assert W.rank() == 1
    assert _rel(W.to_dense(), L3 @ S @ L3.T) < 1e-13
    with pytest.raises(ValueError):      # rank-0 input: the reference fails in maximum(abs, ...) of an empty collection
        api.compress_(api.lowrank(api.DeviceMatrix.from_host(L)).zero())


def test_orth_of_zeros_is_empty(emulated):
    """test/runtests.jl:12-19 -- orth(zeros(n, 1)) has no columns (the building block of the Projection shifts)."""
    n, *_ = emulated
    Q = api.orth_restrict([api.DeviceMatrix.from_host(np.zeros((n, 1)))], api.PencilCombo(0.0, 1.0),
                          api.PencilCombo(1.0, 0.0))
    Ep, Ap = Q
    assert Ep.shape == (0, 0) and Ap.shape == (0, 0)


@pytest.mark.parametrize("core", ["definite", "scaled", "indefinite"])
def test_gale_residual_matches_dense(emulated, core):
    """test/residual.jl:7-52 -- residual(prob, zero) is a copy of C; norm(residual(prob, X)) agrees with the dense
    evaluation A'XE + E'XA + C for definite, scaled and indefinite cores."""
    n, E, A, B, C = emulated
    rng = np.random.default_rng(2)
    G = rng.standard_normal((n, 3))
    Cm = api.lowrank(api.DeviceMatrix.from_host(G), np.diag([1.0, 2.0, 0.5]))
    prob = api.GALEProblem(api.PencilCombo(0.0, 1.0), api.PencilCombo(1.0, 0.0), Cm)
    R0 = api.residual(prob, Cm.zero())
    assert R0.Ls[0] is not Cm.Ls[0]                                  # a copy (:7-16)
    assert _rel(R0.to_dense(), Cm.to_dense()) < 1e-14
    L = rng.standard_normal((n, 4)) * 1e-3
    D = {"definite": np.eye(4), "scaled": 3.0 * np.eye(4), "indefinite": np.diag([1.0, -1.0, 2.0, -0.5])}[core]
    X = api.lowrank(api.DeviceMatrix.from_host(L), D)
    Xd = L @ D @ L.T
    dense = A.T @ Xd @ E + E.T @ Xd @ A + Cm.to_dense()
    R = api.residual(prob, X)
    assert abs(api.norm(R) - np.linalg.norm(dense)) <= 1e-10 * np.linalg.norm(dense)
    assert _rel(R.to_dense(), dense) < 1e-10


def test_lowrank_update_smw_solve(emulated):
    """test/LowRankUpdate.jl:20-51 -- (A + inv(alpha) U V) X = B through Sherman-Morrison-Woodbury, for a one-column
    right-hand side and a block; destructuring returns the operands."""
    n, E, A, B, C = emulated
    rng = np.random.default_rng(3)
    U = rng.standard_normal((n, 2)) * 0.1
    V = rng.standard_normal((2, n)) * 0.1
    Ud = api.DeviceMatrix.from_host(U)
    F = api.lr_update(api.PencilCombo(1.0, -0.01), -1.0, Ud, V)
    assert F.U is Ud and F.alpha == -1.0
    M = (A - 0.01 * E).toarray() - U @ V
    for cols in (1, 5):
        Bh = rng.standard_normal((n, cols))
        X = api.solve_block(api.BlockLinearProblem(F, api.DeviceMatrix.from_host(Bh))).to_host()
        assert _rel(M @ X, Bh) < 1e-10
    # plain sparse operator (no update): Backslash path
    X = api.solve_block(api.BlockLinearProblem(api.PencilCombo(1.0, -0.01), api.DeviceMatrix.from_host(Bh))).to_host()
    assert _rel((A - 0.01 * E) @ X, Bh) < 1e-10


# ---- test/Shifts.jl: helper semantics of the product's host mirror (no device needed) ---------------------------------
def test_shift_helpers_semantics():
    """test/Shifts.jl:22-68,98-163,185-226 -- Projection history must be even; Cyclic cycles through fixed values
    (and through an inner strategy's values); Wrapped applies its function to every refill; BufferedIterator takes
    one shift at a time and never refills for peeks; safe_sort keeps conjugate pairs adjacent; unstable Ritz values
    are discarded, or all flipped."""
    with pytest.raises(ValueError):
        api.Projection(1)
    assert api.Projection(2).n_history == 2
    cyc = api.shifts_init(api.Cyclic([-1.0, -2.0, -3.0]), None)
    assert [cyc.take() for _ in range(7)] == [-1.0, -2.0, -3.0, -1.0, -2.0, -3.0, -1.0]
    assert cyc.peek_many(2) == [-2.0, -3.0] and cyc.take() == -2.0        # peeking consumes nothing

    class Fixed(api.Shifts.Strategy):
        def init(self, prob):
            return api._ListIterator([-4.0, -5.0])

    cyc2 = api.shifts_init(api.Cyclic(Fixed()), None)
    assert [cyc2.take() for _ in range(5)] == [-4.0, -5.0, -4.0, -5.0, -4.0]

    calls = []

    class Gen:
        def __init__(self):
            self.k = 0

        def update(self, *args):
            calls.append(("update", len(args)))

        def take_many(self):
            self.k += 1
            return [-1.0 * self.k, -10.0 * self.k]

    class GenStrategy(api.Shifts.Strategy):
        def init(self, prob):
            return api.BufferedIterator(Gen())

    buf = api.shifts_init(GenStrategy(), None)
    assert buf.peek_many(3) == []                                           # never triggers take_many!
    assert [buf.take() for _ in range(3)] == [-1.0, -10.0, -2.0]
    assert buf.peek_many(3) == [-20.0]
    buf.update("X", "R", "V")
    assert calls == [("update", 3)]
    wrapped = api.shifts_init(api.Wrapped(lambda v: [2 * x for x in v], GenStrategy()), None)
    assert isinstance(wrapped, api.BufferedIterator)                        # helpers.jl:96-99: stays buffered
    assert [wrapped.take() for _ in range(3)] == [-2.0, -20.0, -4.0]
    # safe_sort: by real part, conjugate pairs adjacent with the same order of their members (helpers.jl:122)
    s = api.Shifts.safe_sort([-1 + 2j, -3.0, -1 - 2j, -2 + 1j, -2 - 1j])
    assert s[0] == -3.0 and {s[1], s[2]} == {-2 + 1j, -2 - 1j} and {s[3], s[4]} == {-1 + 2j, -1 - 2j}
    with pytest.warns(UserWarning):
        assert api.Shifts.stabilize_ritz_values([-1.0, 2.0, -3.0], "x") == [-1.0, -3.0]
    with pytest.warns(UserWarning):
        assert api.Shifts.stabilize_ritz_values([1.0, 2.0], "x") == [-1.0, -2.0]
    assert api.Shifts.stabilize_ritz_values([-1.0, -2.0], "x") == [-1.0, -2.0]
    # Penzl heuristic on a real spectrum returns the requested number of shifts from the candidates
    P = api.Shifts.heuristic([-1.0, -10.0, -100.0, -1000.0], 3)
    assert len(P) == 3 and set(P) <= {-1.0, -10.0, -100.0, -1000.0}
