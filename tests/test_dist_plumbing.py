"""CPU tests (gloo, world_size 2) of the multi-GPU plumbing in dre_b200.dist: column partition, all-gather
of column blocks, agreement broadcasts.  The CUDA side of the sharded solve needs GPUs
(tests/test_gpu_dist.py, run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

import dre_b200  # noqa: F401
from dre_b200 import dist as ddist


def test_partition_covers_all_columns():
    for r in (0, 1, 7, 242, 250, 256):
        for w in (1, 2, 3, 4, 8):
            blocks = ddist.partition(r, w)
            assert len(blocks) == w
            assert blocks[0][0] == 0 and blocks[-1][1] == r
            assert all(a <= b for a, b in blocks)
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            widths = [b - a for a, b in blocks]
            assert max(widths) - min(w_ for w_ in widths if w_ > 0 or True) <= -(-r // w)  # ceil blocks, ragged tail


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ncols, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ddist.enable()
        assert ddist.active() and ddist.state().rank == rank
        n = 37
        full = torch.arange(n * ncols, dtype=torch.float64).reshape(n, ncols) * 0.5 - 3.0
        blocks = ddist.partition(ncols, world)
        c0, c1 = blocks[rank]
        got = ddist.allgather_columns(full[:, c0:c1].contiguous(), [b - a for a, b in blocks])
        ok = bool(torch.equal(got, full))
        # control-flow agreement: every rank ends up with rank 0's values
        assert ddist.agree_scalar(3.0 + rank) == 3.0 + rank   # default: trusted (ranks are bit-identical)
        ddist._PARANOID = True
        x = ddist.agree_scalar(1.25 if rank == 0 else 99.0)
        arr = ddist.agree_array(np.array([1 + 2j, -3.5j]) if rank == 0 else np.array([7.0]))
        ok = ok and x == 1.25 and np.array_equal(arr, np.array([1 + 2j, -3.5j]))
        ddist.assert_same_int(42, "test value")
        raised = False
        try:
            ddist.assert_same_int(rank, "deliberately different")
        except RuntimeError:
            raised = True
        q.put((rank, ok and raised))
    finally:
        ddist.disable()
        tdist.destroy_process_group()


@pytest.mark.parametrize("ncols", [10, 7, 1])
def test_allgather_and_agreement_world2(ncols):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ncols, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
