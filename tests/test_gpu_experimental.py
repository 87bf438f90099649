"""GPU tests of the opt-in code paths (environment switches, default off).  They were developed on the host-side SIMT
emulator (tests/test_simt_kernels.py) and measured on the B200 in round 2 (profiles/r02_results.md: A/B of every
switch on the headline workload): the row-split sweeps (DRE_SWEEP2), the narrow k_diag (DRE_DIAG_NARROW_MIN), k_spmm2
(DRE_SPMM2), the eager remainder projection (DRE_RR_EAGER) and the overlapped residual norm (DRE_ASYNC_NORM) did not
beat the defaults and stay off; the compression lane (DRE_ASYNC_COMPRESS) gained 6-14 % on one GPU and is superseded
by the two-GPU pipeline mode of dre_b200.dist.  The tests keep the alternative kernels parity-checked."""
import os
import warnings

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import dre_b200
from dre_b200 import api

pytestmark = [pytest.mark.gpu]
pencils = dre_b200.pencils


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture()
def sweep2(monkeypatch):
    """A fresh context with the row-split sweeps (the flag is read when the context is created)."""
    monkeypatch.setenv("DRE_SWEEP2", "1")
    api.reset_backend()
    yield
    api.reset_backend()


@pytest.mark.parametrize("n,leaf,cap", [(1357, 96, 256), (5177, 24, 40), (20209, 96, 256)])
def test_row_split_sweeps_against_superlu(sweep2, monkeypatch, n, leaf, cap):
    """DRE_SWEEP2=1: M21 panels + k_fwd2 / k_bwd2, 250 / 37 / 3 right-hand sides, real and complex shift,
    plain and closed-loop (SMW) operator."""
    monkeypatch.setenv("DRE_LEAF_SIZE", str(leaf))
    monkeypatch.setenv("DRE_MAX_SNODE", str(cap))
    E, A, B, C, _ = pencils.rail_pencil(n)
    api.upload_pencil(E, A)
    rng = np.random.default_rng(3)
    a, e = 1.0, -1.0 / 200.0
    F = api.PencilCombo(a, e)
    for mu in (-0.37, -0.02 + 0.11j):
        cx = isinstance(mu, complex)
        dtype = complex if cx else float
        M = (a * A + (e + mu) * E).tocsc().astype(dtype)
        lu = spla.splu(M)
        for nrhs in (250, 37, 3):
            R = rng.standard_normal((n, nrhs))
            out = api.solve_block(api.BlockLinearProblem(F, api.DeviceMatrix.from_host(R)), mu=mu)
            V = out[0].to_host() + 1j * out[1].to_host() if cx else out.to_host()
            assert _rel(V, lu.solve(R.astype(dtype))) < 1e-10, (mu, nrhs)
            assert _rel(M @ V, R) < 1e-11, (mu, nrhs)


@pytest.mark.parametrize("n,nsteps,ros", [(371, 3, 1), (371, 2, 2), (1357, 2, 1)])
def test_row_split_sweeps_lockstep_parity(sweep2, n, nsteps, ros):
    """The parity gate of tests/test_gpu_parity.py (K(t) 1e-8, residual norms 1e-10, equal iteration counts, shifts
    replayed from the oracle) with the row-split sweeps."""
    from tests import test_gpu_parity as P
    from oracle import dre_oracle as O

    dt = -100.0 if ros == 1 else -50.0
    so, ro = P._oracle_run(n, nsteps, O.Ros1() if ros == 1 else O.Ros2(), dt=dt)
    adi = api.ADI(shifts=P.ForcedShifts([r["shifts"] for r in ro.runs]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sg, rg = P._gpu_run(n, nsteps, api.Ros1(adi) if ros == 1 else api.Ros2(adi), dt=dt)
    for Ko, Kg in zip(so.K, sg.K):
        assert np.linalg.norm(Kg - Ko) <= 1e-8 * np.linalg.norm(Ko)
    assert [r["iters"] for r in ro.runs] == [r["iters"] for r in rg.runs]
    for a_, b_ in zip(ro.runs, rg.runs):
        ra = np.array([x for _, x in a_["res"]])
        rb = np.array([x for _, x in b_["res"]])
        assert np.max(np.abs(ra - rb) / ra) <= 1e-10


def test_spmm2_and_narrow_diag(monkeypatch):
    """DRE_SPMM2=1 (k_spmm2) against SciPy, and DRE_DIAG_NARROW_MIN=1 (64-thread k_diag on every level): the block
    solve must not change by a single bit (both knobs are re-read whenever a context is created)."""
    n = 5177
    E, A, B, C, _ = pencils.rail_pencil(n)
    rng = np.random.default_rng(4)
    R = rng.standard_normal((n, 250))
    F = api.PencilCombo(1.0, -1.0 / 200.0)
    M = (A + (-1.0 / 200.0 - 0.37) * E).tocsc()
    outs = []
    for narrow in (None, "1"):
        if narrow:
            monkeypatch.setenv("DRE_DIAG_NARROW_MIN", narrow)
        monkeypatch.setenv("DRE_SPMM2", "1")
        api.reset_backend()
        api.upload_pencil(E, A)
        X = api.DeviceMatrix.from_host(R)
        Y = api.spmm("E", X, alpha=-0.7, Y=api.DeviceMatrix.from_host(R), beta=1.0)
        assert _rel(Y.to_host(), R - 0.7 * (E @ R)) < 1e-13
        out = api.solve_block(api.BlockLinearProblem(F, X), mu=-0.37).to_host()
        assert _rel(M @ out, R) < 1e-11
        outs.append(out)
    assert np.array_equal(outs[0], outs[1])
    monkeypatch.delenv("DRE_SPMM2")
    monkeypatch.delenv("DRE_DIAG_NARROW_MIN")
    api.reset_backend()


@pytest.mark.parametrize("seed", [0, 1])
def test_gmres_and_fgmres_tiny_random(seed):
    """SURVEY 8f rank 2 (test/tiny_random.jl:25-46): low-rank GMRES and FGMRES with an ADI preconditioner on the GPU
    against dense Bartels-Stewart (developed on the emulator: tests/test_simt_endtoend.py)."""
    from oracle import dre_oracle as O

    api.reset_backend()
    n, g = 50, 4
    rng = np.random.default_rng(seed)
    E, A = pencils.random_spd_pencil(n, seed=seed)
    G = rng.random((n, g))
    prob_o = O.GALEProblem(E, A, -2 * O.lowrank(G, -np.eye(g)))
    res0 = O.norm(prob_o.C)
    X_ref = O.bartels_stewart(prob_o)

    def run(alg):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return api.solve(api.GALEProblem(E, A, -2 * api.lowrank(G, -np.eye(g))), alg).to_dense()

    X_gmres = run(api.GMRES(maxiters=5, reltol=1e-8))
    X_fgmres = run(api.GMRES(maxiters=3, maxrestarts=0, reltol=1e-10, preconditioner=api.ADI(
        maxiters=10, shifts=api.Cyclic(api.Heuristic(10, 10, 10)), compression_interval=20, warn_convergence=False)))
    assert np.linalg.norm(O.gale_residual_dense(prob_o, X_gmres)) / res0 < 2e-8
    assert np.linalg.norm(O.gale_residual_dense(prob_o, X_fgmres)) / res0 < 1e-10
    assert O.delta(X_gmres, X_ref) < 2e-8
    assert O.delta(X_fgmres, X_ref) < 1e-10
    api.reset_backend()


@pytest.mark.parametrize("n,nsteps,ros", [(371, 3, 1), (1357, 2, 1), (371, 2, 2)])
def test_async_norm_lockstep_parity(monkeypatch, n, nsteps, ros):
    """DRE_ASYNC_NORM (side-stream residual norm + speculative solve of the next shift): the lock-step parity gate,
    and the speculation must really be adopted."""
    from tests import test_gpu_parity as P
    from oracle import dre_oracle as O

    api.reset_backend()
    monkeypatch.setattr(api, "ASYNC_NORM", True)
    adopted = []
    orig = api.solve_

    def counting_solve(cache):
        out = orig(cache)
        adopted.append((getattr(cache, "adopted_speculations", 0), len(cache.shifts)))
        return out

    monkeypatch.setattr(api, "solve_", counting_solve)
    dt = -100.0 if ros == 1 else -50.0
    so, ro = P._oracle_run(n, nsteps, O.Ros1() if ros == 1 else O.Ros2(), dt=dt)
    adi = api.ADI(shifts=P.ForcedShifts([r["shifts"] for r in ro.runs]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sg, rg = P._gpu_run(n, nsteps, api.Ros1(adi) if ros == 1 else api.Ros2(adi), dt=dt)
    for Ko, Kg in zip(so.K, sg.K):
        assert np.linalg.norm(Kg - Ko) <= 1e-8 * np.linalg.norm(Ko)
    assert [r["iters"] for r in ro.runs] == [r["iters"] for r in rg.runs]
    for a_, b_ in zip(ro.runs, rg.runs):
        ra = np.array([x for _, x in a_["res"]])
        rb = np.array([x for _, x in b_["res"]])
        assert np.max(np.abs(ra - rb) / ra) <= 1e-10
    assert adopted[0][0] >= 0.4 * adopted[0][1], adopted
    api.reset_backend()


@pytest.mark.parametrize("n,nsteps", [(371, 3), (1357, 2)])
def test_async_compress_lane_lockstep_parity(monkeypatch, n, nsteps):
    """DRE_ASYNC_COMPRESS (compression lane: second context + host thread) together with DRE_ASYNC_NORM: the
    lock-step parity gate must hold unchanged."""
    from tests import test_gpu_parity as P
    from oracle import dre_oracle as O

    api.reset_backend()
    monkeypatch.setattr(api, "ASYNC_COMPRESS", True)
    monkeypatch.setattr(api, "ASYNC_NORM", True)
    so, ro = P._oracle_run(n, nsteps, O.Ros1(), dt=-100.0)
    adi = api.ADI(shifts=P.ForcedShifts([r["shifts"] for r in ro.runs]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sg, rg = P._gpu_run(n, nsteps, api.Ros1(adi), dt=-100.0)
    for Ko, Kg in zip(so.K, sg.K):
        assert np.linalg.norm(Kg - Ko) <= 1e-8 * np.linalg.norm(Ko)
    assert [r["iters"] for r in ro.runs] == [r["iters"] for r in rg.runs]
    for a_, b_ in zip(ro.runs, rg.runs):
        ra = np.array([x for _, x in a_["res"]])
        rb = np.array([x for _, x in b_["res"]])
        assert np.max(np.abs(ra - rb) / ra) <= 1e-10
    api.reset_backend()
