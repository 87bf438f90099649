"""GPU parity at the sizes BASELINE.json quotes (pytest -m gpu), through the C ABI, against the CPU oracle.

* config 4 (n = 79 841, low-rank Ros1): lock-step (the GPU ADI replays the shifts the oracle consumed) over the
  first time step in full and a second step with ADI(maxiters=25) on both sides (the oracle needs ~6 s per ADI
  iteration at this size);
* config 3 (n = 20 209, low-rank Ros2, complex shift pairs): lock-step over two time steps with
  ADI(maxiters=25) on both sides (four ADI solves, real and complex shifted factorizations);
* free run (the GPU generates its own Projection(2) shifts): at every refill of the shift buffer the oracle's
  orth() (src/Stuff.jl:13-18: keep singular values > n*eps, ABSOLUTE) is evaluated on the very V blocks the GPU
  run holds.  Wherever the GPU path and the oracle formula keep a different number of directions, the deciding
  singular values must lie within 4x of n*eps -- the knife edge of the reference's own threshold
  (test/cuda.jl:86-100 is the template: accelerated vs CPU K(t) at every step).

Tolerances (BASELINE.json north_star): K(t) <= 1e-8 relative at every saved time point, ADI residual norms within
1e-10 relative, identical ADI iteration counts.
"""
import os
import warnings

import numpy as np
import pytest
import scipy.linalg as sla

from dre_b200 import api
from oracle import dre_oracle as O

from .test_gpu_parity import ForcedShifts, Recorder, _problem

pytestmark = pytest.mark.gpu


def _blas_threads():
    """The big-n oracle runs are dominated by LAPACK QR/eig of n x ~3000 panels: let them use the host cores
    (tests/conftest.py pins BLAS to one thread for the many tiny problems of the other tests)."""
    try:
        from threadpoolctl import threadpool_limits

        return threadpool_limits(limits=min(16, os.cpu_count() or 1))
    except Exception:  # pragma: no cover
        import contextlib

        return contextlib.nullcontext()


def _complexify(shifts):
    """Shifts.Wrapped function (src/shifts/helpers.jl:48-58) that turns every other adjacent pair of REAL Ritz
    values (a, b) into the conjugate pair sqrt(ab) exp(+-0.3i): at n = 20 209 the Projection(2) Ritz values of the
    first time steps are all real, and the complex double step (adi.jl:181-225) is what config 3 is about."""
    out, arr, i = [], list(shifts), 0
    while i < len(arr):
        a = arr[i]
        b = arr[i + 1] if i + 1 < len(arr) else None
        if b is not None and np.imag(a) == 0 and np.imag(b) == 0 and (i // 2) % 2 == 0:
            z = -np.sqrt(float(np.real(a)) * float(np.real(b))) * np.exp(0.3j)
            out += [z, np.conj(z)]
            i += 2
        elif b is not None and np.imag(a) != 0:   # an existing conjugate pair stays adjacent
            out += [a, b]
            i += 2
        else:
            out.append(a)
            i += 1
    return np.array(out)


def _lockstep(n, nsteps, ros, dt, maxiters, wrap=None):
    E, A, B, C, L0, D0 = _problem(n)
    tspan = (4500.0, 4500.0 + nsteps * dt)
    ro, rg = Recorder(), Recorder()
    with warnings.catch_warnings(), _blas_threads():
        warnings.simplefilter("ignore")
        shifts_o = O.Wrapped(wrap, O.Projection(2)) if wrap is not None else None
        alg_o = (O.Ros1 if ros == 1 else O.Ros2)(O.ADI(maxiters=maxiters, shifts=shifts_o))
        so = O.solve_gdre(O.GDREProblem(E, A, B, C, O.lowrank(L0, D0), tspan), alg_o, dt=dt, observer=ro)
        adi = api.ADI(maxiters=maxiters, shifts=ForcedShifts([r["shifts"] for r in ro.runs]))
        alg_g = (api.Ros1 if ros == 1 else api.Ros2)(adi)
        sg = api.solve(api.GDREProblem(E, A, B, C, api.lowrank(L0, D0), tspan), alg_g, dt=dt, observer=rg)
    assert len(so.K) == len(sg.K) == nsteps + 1 and so.t == sg.t
    kerr = [float(np.linalg.norm(Kg - Ko) / np.linalg.norm(Ko)) for Ko, Kg in zip(so.K, sg.K)]
    assert max(kerr) <= 1e-8, kerr
    assert [r["iters"] for r in ro.runs] == [r["iters"] for r in rg.runs]
    rerr = 0.0
    for a, b in zip(ro.runs, rg.runs):
        assert [i for i, _ in a["res"]] == [i for i, _ in b["res"]]
        ra = np.array([x for _, x in a["res"]])
        rb = np.array([x for _, x in b["res"]])
        rerr = max(rerr, float(np.max(np.abs(ra - rb) / ra)))
    assert rerr <= 1e-10, rerr
    print(f"lock-step n={n} Ros{ros}: K(t) rel err {kerr}, ADI residual norms rel err {rerr:.2e}, "
          f"iterations {[r['iters'] for r in rg.runs]}")
    return ro, rg


def test_lockstep_ros1_n79841_headline_config():
    """BASELINE config 4 / the bench workload: first step in full, second step capped at 25 ADI iterations."""
    E, A, B, C, L0, D0 = _problem(79841)
    dt = -100.0
    ro, rg = Recorder(), Recorder()
    with warnings.catch_warnings(), _blas_threads():
        warnings.simplefilter("ignore")
        # step 1 in full on both sides; its final X is the initial value of the capped second step
        p1 = (4500.0, 4400.0)
        so1 = O.solve_gdre(O.GDREProblem(E, A, B, C, O.lowrank(L0, D0), p1), O.Ros1(), dt=dt, observer=ro,
                           save_state=True)
        adi = api.ADI(shifts=ForcedShifts([r["shifts"] for r in ro.runs]))
        sg1 = api.solve(api.GDREProblem(E, A, B, C, api.lowrank(L0, D0), p1), api.Ros1(adi), dt=dt, observer=rg,
                        save_state=True)
        n1 = len(ro.runs)
        p2 = (4400.0, 4300.0)
        so2 = O.solve_gdre(O.GDREProblem(E, A, B, C, so1.X[-1], p2), O.Ros1(O.ADI(maxiters=25)), dt=dt, observer=ro)
        adi2 = api.ADI(maxiters=25, shifts=ForcedShifts([r["shifts"] for r in ro.runs[n1:]]))
        sg2 = api.solve(api.GDREProblem(E, A, B, C, sg1.X[-1], p2), api.Ros1(adi2), dt=dt, observer=rg)
    kerr = [float(np.linalg.norm(Kg - Ko) / np.linalg.norm(Ko)) for Ko, Kg in zip(so1.K + so2.K, sg1.K + sg2.K)]
    assert max(kerr) <= 1e-8, kerr
    assert [r["iters"] for r in ro.runs] == [r["iters"] for r in rg.runs]
    rerr = 0.0
    for a, b in zip(ro.runs, rg.runs):
        ra = np.array([x for _, x in a["res"]])
        rb = np.array([x for _, x in b["res"]])
        assert ra.shape == rb.shape
        rerr = max(rerr, float(np.max(np.abs(ra - rb) / ra)))
    assert rerr <= 1e-10, rerr
    print(f"lock-step n=79841 Ros1: K(t) rel err {kerr}, ADI residual norms rel err {rerr:.2e}, "
          f"iterations {[r['iters'] for r in rg.runs]}")


def test_lockstep_ros2_n20209_config3():
    """BASELINE config 3: low-rank Ros2 at n = 20 209, complex shift pairs asserted (half of the Projection(2)
    Ritz pairs are rotated off the real axis by a Shifts.Wrapped function on the oracle side; the GPU replays)."""
    ro, rg = _lockstep(20209, 2, 2, -50.0, 25, wrap=_complexify)
    ncomplex = sum(1 for r in rg.runs for s_ in r["shifts"] if s_.imag != 0)
    assert ncomplex >= 20, ncomplex


@pytest.mark.parametrize("n,nsteps", [(371, 3), (5177, 2)])
def test_free_run_divergence_is_the_orth_knife_edge(n, nsteps):
    """Free run.  Every refill of the Projection(2) shift buffer of the GPU run is re-evaluated with the oracle's
    orth() on the same V blocks: the number of kept directions may only differ where the deciding singular values
    sit within 4x of the reference's absolute threshold n*eps (src/Stuff.jl:15-16).  K(t) must agree with the
    oracle's own free run to max(1e-8, final relative ADI residual of the steps before it): two incomplete ADI
    solves of the same Lyapunov equation differ by O(residual)."""
    E, A, B, C, L0, D0 = _problem(n)
    tspan = (4500.0, 4500.0 - 100.0 * nsteps)
    thr = n * O.EPS
    log = []
    orig = api.orth_restrict

    def traced(Vs, Ep, Ap):
        api.ORTH_TRACE = []
        out = orig(Vs, Ep, Ap)
        tr = api.ORTH_TRACE[-1] if api.ORTH_TRACE else None
        api.ORTH_TRACE = None
        N = np.concatenate([V.to_host() for V in Vs], axis=1)
        s = sla.svdvals(N)
        log.append(dict(k=N.shape[1], kept_oracle=int(np.count_nonzero(s > thr)),
                        kept_gpu=(tr["kept"] if tr else 0), s=s))
        return out

    ro, rg = Recorder(), Recorder()
    api.orth_restrict = traced
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sg = api.solve(api.GDREProblem(E, A, B, C, api.lowrank(L0, D0), tspan), api.Ros1(), dt=-100.0, observer=rg)
    finally:
        api.orth_restrict = orig
        api.ORTH_TRACE = None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        so = O.solve_gdre(O.GDREProblem(E, A, B, C, O.lowrank(L0, D0), tspan), O.Ros1(), dt=-100.0, observer=ro)
    assert log, "the free run never refilled its shift buffer"
    ndiff, first = 0, None
    for i, e in enumerate(log):
        lo, hi = sorted((e["kept_oracle"], e["kept_gpu"]))
        if lo == hi:
            continue
        ndiff += 1
        deciding = e["s"][lo:hi]
        if first is None:
            first = (i, e["k"], e["kept_oracle"], e["kept_gpu"], deciding / thr)
        assert np.all(deciding <= 4 * thr) and np.all(deciding >= thr / 4), (i, e["k"], lo, hi, deciding / thr)
    print(f"free run n={n}: {len(log)} refills, {ndiff} with a different number of kept directions; first: {first}; "
          f"iterations oracle {[r['iters'] for r in ro.runs]} gpu {[r['iters'] for r in rg.runs]}")
    for i, (Ko, Kg) in enumerate(zip(so.K, sg.K)):
        unconverged = max([r["res"][-1][1] / r["res"][0][1] for r in ro.runs[:i]] +
                          [r["res"][-1][1] / r["res"][0][1] for r in rg.runs[:i]] + [0.0])
        err = np.linalg.norm(Kg - Ko) / np.linalg.norm(Ko)
        assert err <= max(1e-8, 10 * unconverged), (i, err, unconverged)
    if ndiff == 0:
        # no knife edge met: the shift sequences coincide up to the conditioning of the (nonsymmetric) Ritz problems,
        # so the iteration counts of converging solves may move by a step or two but not more
        for a, b in zip(ro.runs, rg.runs):
            assert abs(a["iters"] - b["iters"]) <= max(2, 0.1 * a["iters"]), (a["iters"], b["iters"])


def test_config5_heat3d_gale_adi_and_newton_small():
    """BASELINE config 5 at test size (3D heat pencil 14^3 = 2744 unknowns, 8 inputs / 8 outputs; the bench line
    `bench.py --config 5` runs 60^3 and larger): standalone GALE ADI with a dense identity core
    (test/tiny_random.jl:15-19) against the dense Lyapunov residual, lock-step against the oracle's ADI, and
    Newton-ADI for the GARE (test/rail.jl:74-88: residual < reltol ||Q||)."""
    import dre_b200

    E, A, B, C, _ = dre_b200.pencils.heat3d_pencil(14)
    q = C.shape[0]
    ro, rg = Recorder(), Recorder()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Co = O.lowrank(np.asfortranarray(C.T), np.eye(q))
        prob_o = O.GALEProblem(E, A, Co)
        Xo = O.solve_gale(prob_o, O.ADI(), observer=ro)
        Cg = api.lowrank(np.asfortranarray(C.T), np.eye(q))
        adi = api.ADI(shifts=ForcedShifts([r["shifts"] for r in ro.runs]))
        Xg = api.solve(api.GALEProblem(E, A, Cg), adi, observer=rg)
    assert [r["iters"] for r in ro.runs] == [r["iters"] for r in rg.runs]
    ra = np.array([x for _, x in ro.runs[0]["res"]])
    rb = np.array([x for _, x in rg.runs[0]["res"]])
    assert np.max(np.abs(ra - rb)) <= 1e-10 * O.norm(Co)
    Xd = Xg.to_dense()
    assert np.linalg.norm(O.gale_residual_dense(prob_o, Xd)) <= 1e-10 * O.norm(Co)
    assert O.delta(Xd, Xo.to_dense()) <= 1e-9
    reltol = 1e-10
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        are = api.GAREProblem(E, A, api.lowrank(B), api.lowrank(np.asfortranarray(C.T)))
        Xn = api.solve(are, api.Newton(api.ADI(ignore_initial_guess=True), maxiters=10, reltol=reltol))
    are_o = O.GAREProblem(E, A, O.lowrank(B), O.lowrank(np.asfortranarray(C.T)))
    assert np.linalg.norm(O.gare_residual_dense(are_o, Xn.to_dense())) < reltol * O.norm(are_o.Q)
