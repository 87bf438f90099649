"""GPU tier, collected last (file name sorts after test_zz_golden.py): tests written after the round's last GPU call and
so far exercised only on the CPU emulator tier (their emulator twins are green).  Kept at the end so that the
parity gates of the other files are recorded first under `pytest -x`."""
import numpy as np
import pytest

import dre_b200  # noqa: F401
from dre_b200 import api, pencils

pytestmark = pytest.mark.gpu

_RAIL = pencils.rail_pencil(1357)[:4]


@pytest.fixture()
def rail():
    E, A, B, C = _RAIL
    api.upload_pencil(E, A)  # no-op while the same pencil is resident
    return E, A, B, C


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _streamed_compress_case(n, rng):
    """An ADI-shaped compress!: an orthonormal first term with a diagonal core plus increments that mostly lie in its
    span, one term with a dense core; returns (terms as host arrays, dense X)."""
    Q0, _ = np.linalg.qr(rng.standard_normal((n, 40)))
    d0 = rng.standard_normal(40)
    terms = [(1.0, Q0, np.diag(d0))]
    dense = Q0 @ np.diag(d0) @ Q0.T
    for i in range(4):
        V = Q0 @ rng.standard_normal((40, 24)) * 10.0 ** (-i) + 10.0 ** (-2 * i - 1) * rng.standard_normal((n, 24))
        if i == 2:
            S = rng.standard_normal((24, 24))
            S = S + S.T
        else:
            S = np.diag(rng.standard_normal(24))
        a = -0.7 * (i + 1)
        terms.append((a, V, S))
        dense = dense + a * V @ S @ V.T
    return terms, dense


def test_streamed_compress_matches_one_call(rail):
    """dre_compress_begin / _add / _finish (the compression lane of the multi-GPU pipeline mode adds the terms as they
    arrive) against dre_ldlt_compress on the same terms, and both against the dense sum (src/LDLt.jl:204-225)."""
    E, A, B, C = rail
    n = E.shape[0]
    terms, dense = _streamed_compress_case(n, np.random.default_rng(11))
    be = api.backend()
    dev = [(a, api.DeviceMatrix.from_host(L), np.asfortranarray(D)) for a, L, D in terms]
    L1, lam1 = api._compress_call(be, dev)
    job = api.CompressStream(be, sum(L.shape[1] for _, L, _ in terms) + 64)
    job.add(dev[:2])
    for t in dev[2:]:
        assert job.room_for(t[1].ncols)
        job.add([t])
    L2, lam2 = job.finish()
    X1 = L1.to_host() @ np.diag(lam1) @ L1.to_host().T
    X2 = L2.to_host() @ np.diag(lam2) @ L2.to_host().T
    assert _rel(X1, dense) < 1e-12 and _rel(X2, dense) < 1e-12
    assert abs(len(lam1) - len(lam2)) <= 2   # (two eigenvalues of this case sit within 2x of the truncation threshold)
    assert np.linalg.norm(L2.to_host().T @ L2.to_host() - np.eye(len(lam2))) < 1e-11
    # a job that runs out of room reports it instead of overrunning its workspace
    job = api.CompressStream(be, 50)
    job.add(dev[:1])
    assert not job.room_for(24)
    with pytest.raises(Exception):
        job.add(dev[1:2])
    # while a job is open the rank-revealing QR (shared workspaces) refuses to run; finishing the job releases it
    with pytest.raises(Exception):
        api.orth_restrict([dev[1][1]], api.PencilCombo(0.0, 1.0), api.PencilCombo(1.0, 0.0))
    L3, lam3 = job.finish()
    assert _rel(L3.to_host() @ np.diag(lam3) @ L3.to_host().T, terms[0][1] @ terms[0][2] @ terms[0][1].T) < 1e-12
    api.orth_restrict([dev[1][1]], api.PencilCombo(0.0, 1.0), api.PencilCombo(1.0, 0.0))



def test_pencil_is_recognised_by_content(rail):
    """an equal copy of the resident pencil does not trigger a re-upload (which would invalidate every
    DeviceMatrix), values changed in place do."""
    E, A, B, C = rail
    be = api.backend()
    g = be.generation
    M = api.DeviceMatrix.from_host(np.ones((E.shape[0], 2)))
    E2 = E.copy()
    api.upload_pencil(E2, A)
    assert be.generation == g
    assert np.array_equal(M.to_host(), np.ones((E.shape[0], 2)))      # the panel is still valid
    E2.data *= 2.0
    api.upload_pencil(E2, A)
    assert be.generation == g + 1
    Y = api.spmm("E", api.DeviceMatrix.from_host(np.ones((E.shape[0], 2))))
    assert _rel(Y.to_host(), 2.0 * (E @ np.ones((E.shape[0], 2)))) < 1e-14
    api.upload_pencil(E, A)
    assert be.generation == g + 2
