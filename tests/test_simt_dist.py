"""CPU tier, world_size 2 over gloo: the column-sharded ADI path of dre_b200.dist end to end on the
host-side SIMT emulator -- every rank runs the emulated C ABI (tests/simt/), solves its column block with
dre_adi_solve, the blocks are exchanged with the same all-gather code as on the GPUs (torch.distributed, gloo
instead of NCCL), residual update / norms / compression / shifts run replicated.  The sharded solve must reproduce
the single-process K(t) and iteration counts, and the ranks must stay bit-identical.
Only the two torch.cuda-specific helpers of dist.py (raw-pointer tensor view, library stream) are replaced by CPU
equivalents inside the test processes; the product code is unchanged."""
import contextlib
import ctypes as C
import os
import socket
import warnings

import numpy as np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, n, q):
    os.environ["DRE_NO_PRIME"] = "1"
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    import scipy.sparse.linalg as spla
    import torch
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
        tdist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=300))

        def panel_tensor(be, M):   # CPU stand-in for the __cuda_array_interface__ view of library memory
            ptr, ld = C.c_void_p(), C.c_int64()
            be.check(be.lib.dre_mat_devptr(be.h, M.view, C.byref(ptr), C.byref(ld)))
            buf = (C.c_double * (be.n * ld.value)).from_address(ptr.value)
            return torch.from_numpy(np.ctypeslib.as_array(buf).reshape(be.n, ld.value))[:, :M.ncols]

        ddist.panel_tensor = panel_tensor
        ddist._library_stream = lambda be: None
        torch.cuda.stream = lambda s: contextlib.nullcontext()
        ddist.enable(device=None)
    api.backend()
    E, A, B, Cm, _ = dre_b200.pencils.rail_pencil(n)
    L0 = spla.splu(E.tocsc()).solve(Cm.T)
    iters = []

    class Obs:
        def observe_gale_done(self, it, X, res, rn):
            iters.append(it)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = api.solve(api.GDREProblem(E, A, B, Cm, api.lowrank(L0, 0.01 * np.eye(Cm.shape[0])), (4500.0, 4400.0)),
                        api.Ros1(), dt=-100.0, observer=Obs())
    gathered = ddist.state().bytes_gathered if world > 1 else 0
    q.put((rank, [np.asarray(K) for K in sol.K], iters, gathered))
    if world > 1:
        ddist.disable()
        tdist.destroy_process_group()


def _spawn(world, n):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return sorted(res, key=lambda t: t[0])


import pytest


@pytest.mark.parametrize("lane", [False, True])
def test_sharded_ros1_on_the_emulator_matches_single_process(lane, monkeypatch):
    """lane=True: every rank additionally runs compress!(X) on its own compression lane (DRE_ASYNC_COMPRESS)."""
    monkeypatch.setenv("DRE_ASYNC_COMPRESS", "1" if lane else "0")
    n = 371
    single = _spawn(1, n)[0]
    sharded = _spawn(2, n)
    assert sharded[0][3] > 0                      # the all-gather really ran
    for r in sharded:
        assert r[2] == single[2]                  # identical ADI iteration counts on every rank
        for Ks, K1 in zip(r[1], single[1]):
            assert np.linalg.norm(Ks - K1) <= 1e-10 * np.linalg.norm(K1)
    for Ka, Kb in zip(sharded[0][1], sharded[1][1]):
        assert np.array_equal(Ka, Kb)             # ranks stay bit-identical
