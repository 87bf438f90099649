"""CPU tier, world_size 2 and 3 over gloo: the PIPELINE mode of dre_b200.dist end to end on the host-side SIMT
emulator.  Rank 0 runs the low-rank Ros1 driver (ADI iteration chain), streams every increment of X to rank 1,
which holds X and runs compress! there (a third rank, if any, idles); control messages and -- on the CPU -- the
panels travel over gloo, exactly the code path the GPUs use with NCCL for the panels.  The result must reproduce
the single-process K(t) and ADI iteration counts to round-off (same arithmetic on the same operands)."""
import os
import socket
import warnings

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, n, nsteps, ros, q, lanes=1, forced=None):
    os.environ["DRE_NO_PRIME"] = "1"
    os.environ["DRE_PIPE_LANES"] = str(lanes)
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    import scipy.sparse.linalg as spla
    import torch.distributed as tdist

    import dre_b200
    from dre_b200 import api, capi
    from dre_b200 import dist as ddist
    from tests.simt import build_emu

    capi.LIB_PATH, capi._lib = build_emu.build(), None
    if world > 1:
        import datetime

        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        tdist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=600))
        ddist.enable_pipeline(device=None)
        if rank > 0:
            served = ddist.serve(api)
            q.put((rank, served))
            tdist.destroy_process_group()
            return
    api.backend()
    E, A, B, Cm, _ = dre_b200.pencils.rail_pencil(n)
    L0 = spla.splu(E.tocsc()).solve(Cm.T)
    iters, shifts = [], [[]]

    class Obs:
        def observe_gale_metadata(self, desc, mu):
            shifts[-1].append(complex(mu))

        def observe_gale_done(self, it, X, res, rn):
            iters.append(it)
            shifts.append([])

    class Forced(api.Shifts.Strategy):
        """replays, for the i-th ADI solve, the shifts the single-process run consumed in its i-th solve"""

        def __init__(self, lists):
            self.lists, self.i = lists, 0

        def init(self, prob):
            lst = self.lists[self.i]
            self.i += 1
            return api._ListIterator([z.real if z.imag == 0 else z for z in lst] + [-1.0] * 4)

    dt = -100.0 if ros == 1 else -50.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # (30 ADI iterations per solve keep the emulated runs short: three compression points per solve, the second
        #  step starts from the factor the lane returned for the first)
        adi = api.ADI(maxiters=30, shifts=Forced(forced)) if forced is not None else api.ADI(maxiters=30)
        alg = api.Ros1(adi) if ros == 1 else api.Ros2(adi)
        sol = api.solve(api.GDREProblem(E, A, B, Cm, api.lowrank(L0, 0.01 * np.eye(Cm.shape[0])),
                                        (4500.0, 4500.0 + nsteps * dt)), alg, dt=dt, observer=Obs())
    stats = dict(ddist.pipe_state().stats) if world > 1 else {}
    if world > 1:
        ddist.pipe_stop()
    q.put((rank, [np.asarray(K) for K in sol.K], iters, stats, sol.X[-1].rank(), shifts[:len(iters)]))
    if world > 1:
        tdist.destroy_process_group()


_SINGLE = {}


def _single(n, nsteps, ros):
    """the single-process run both tests compare with (once per session)"""
    key = (n, nsteps, ros)
    if key not in _SINGLE:
        _SINGLE[key] = _spawn(1, n, nsteps, ros)[0]
    return _SINGLE[key]


def _spawn(world, n, nsteps, ros, lanes=1, forced=None):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, n, nsteps, ros, q, lanes, forced)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=1500) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return sorted(res, key=lambda t: t[0])


@pytest.mark.parametrize("world,ros", [(3, 1)])
def test_pipeline_mode_matches_single_process(world, ros):
    n, nsteps = 371, 2
    single = _single(n, nsteps, ros)
    piped = _spawn(world, n, nsteps, ros)
    r0 = piped[0]
    assert r0[2] == single[2]                                   # identical ADI iteration counts
    assert r0[4] == single[4]                                   # identical rank of the final X
    for Kp, K1 in zip(r0[1], single[1]):
        assert np.linalg.norm(Kp - K1) <= 1e-10 * np.linalg.norm(K1)
    st = r0[3]
    assert st["terms_sent"] >= sum(single[2]) and st["fetches"] == nsteps and st["compress_cmds"] > 0
    served = piped[1][1]
    assert served["role"] == "compress" and served["terms"] == st["terms_sent"] and served["compressions"] > 0
    if world > 2:
        assert piped[2][1]["role"] == "idle"
    # the second step's initial guess is the factor rank 1 returned for the first one: it is not sent back
    assert st["terms_sent"] == sum(single[2]) + 1               # (+ the initial value X0 of the first step)


def test_two_lane_pipeline_lockstep_with_single_process():
    """DRE_PIPE_LANES=2 on three ranks: the lanes take the compression points in turn, the increments of compress!
    k+1 are orthogonalised before X_k arrives from the other lane and X_k is added last.  A different term order
    means different round-off, so the comparison is the lock-step one (the shifts of the single-process run are
    replayed): K(t) within 1e-8, identical ADI iteration counts, final ranks within one."""
    n, nsteps = 371, 2
    single = _single(n, nsteps, 1)
    piped = _spawn(3, n, nsteps, 1, lanes=2, forced=single[5])
    r0 = piped[0]
    assert r0[2] == single[2]
    assert abs(r0[4] - single[4]) <= 1
    for Kp, K1 in zip(r0[1], single[1]):
        assert np.linalg.norm(Kp - K1) <= 1e-8 * np.linalg.norm(K1)
    st = r0[3]
    assert st["fetches"] == nsteps and st["compress_cmds"] > 0
    lanes = [piped[1][1], piped[2][1]]
    assert all(l["role"] == "compress" for l in lanes)
    assert sum(l["terms"] for l in lanes) == st["terms_sent"]
    assert all(l["compressions"] > 0 and l["handovers"] > 0 for l in lanes)
