"""Multi-GPU layer of the LRSIF-ADI path (BASELINE.json north_star):
one process per GPU (torchrun), every rank holds the pencil, the symbolic analysis and a REPLICATED
per-shift numeric factorization; the right-hand-side columns of every block solve

    V = (F' + mu E')^-1 R                      (src/lyapunov/adi.jl:156-159, :195-198)

are split into contiguous column blocks, one per rank.  The only exchange step of the path is the
all-gather of the solved blocks (NCCL over NVLink, driven through torch.distributed -- plumbing, no
torch type crosses the C ABI: the library exports raw device addresses with ``dre_mat_devptr``);
the residual update R += c E'V, the Gram norms, the column compression and the shift generation run
replicated on identical data, so all ranks stay in lock step.  Scalars that steer control flow
(residual norms, shift lists, compressed ranks) are additionally agreed on through tiny broadcasts /
reductions so that a divergence raises instead of dead-locking a collective.

Nothing here is used unless ``enable()`` was called (single-GPU runs never import torch).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import os

_STATE = None
_PARANOID = os.environ.get("DRE_DIST_PARANOID") is not None   # also broadcast every residual norm from rank 0


class _State:
    def __init__(self, rank, world, group, device):
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.bytes_gathered = 0
        self.gathers = 0


def enable(group=None, device=None):
    """Switch the API mirror to column-sharded block solves.  ``torch.distributed`` must be initialised
    (backend nccl for GPUs; gloo works for the CPU tests of the plumbing)."""
    global _STATE
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("dre_b200.dist.enable: torch.distributed is not initialised")
    _STATE = _State(dist.get_rank(group), dist.get_world_size(group), group, device)
    return _STATE


def disable():
    global _STATE
    _STATE = None


def active() -> bool:
    return _STATE is not None and _STATE.world > 1


def state():
    return _STATE


def partition(ncols: int, world: int):
    """Contiguous column blocks of ceil(ncols / world) columns (the last ranks may get fewer or none)."""
    w = -(-ncols // world) if ncols > 0 else 0
    return [(min(g * w, ncols), min((g + 1) * w, ncols)) for g in range(world)]


def allgather_columns(local, widths, group=None):
    """local: torch tensor [n, widths[rank]] (any device the process group supports).  Returns the
    tensor [n, sum(widths)] whose column blocks are the ranks' ``local`` tensors in rank order."""
    import torch
    import torch.distributed as dist

    world = len(widths)
    n = local.shape[0]
    wmax = max(widths)
    if wmax == 0:
        return local.new_zeros((n, 0))
    pad = local.new_zeros((n, wmax))
    pad[:, :local.shape[1]] = local
    out = local.new_empty((world * n, wmax))   # rank g's block = rows [g*n, (g+1)*n)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[g * n:(g + 1) * n, :widths[g]] for g in range(world)], dim=1)


class _CudaArray:
    """``__cuda_array_interface__`` wrapper of library-owned device memory (row-major n x ncols view)."""

    def __init__(self, ptr, n, ncols, ld):
        self.__cuda_array_interface__ = {"shape": (n, ncols), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": (ld * 8, 8)}


def panel_tensor(be, M):
    """torch view (no copy) of a DeviceMatrix; the caller synchronises the library stream first."""
    import torch

    ptr, ld = C.c_void_p(), C.c_int64()
    be.check(be.lib.dre_mat_devptr(be.h, M.view, C.byref(ptr), C.byref(ld)))
    dev = _STATE.device if _STATE is not None and _STATE.device is not None else 0
    return torch.as_tensor(_CudaArray(ptr.value, be.n, M.ncols, ld.value), device=f"cuda:{dev}")


def _library_stream(be):
    """The context's main stream as a torch ExternalStream: copies and NCCL collectives issued under it are
    stream-ordered with the library's kernels (no host synchronisation on the data path)."""
    import torch

    cached = getattr(be, "_torch_stream", None)
    if cached is not None and cached[0] == be.h.value:
        return cached[1]
    ptr = C.c_void_p()
    be.check(be.lib.dre_get_stream(be.h, C.byref(ptr)))
    dev = _STATE.device if _STATE is not None and _STATE.device is not None else 0
    st = torch.cuda.ExternalStream(ptr.value, device=f"cuda:{dev}")
    be._torch_stream = (be.h.value, st)
    return st


def sharded_adi_solve(be, mu: complex, R, V1, V2, empty_view):
    """Column-sharded version of the solve half of dre_adi_step: this rank solves its block of R's columns,
    then every rank receives all blocks of V1 (and V2 for a complex pair).  Everything is queued on the
    library's stream: solve -> pad/copy of the local block -> NCCL all-gather -> scatter into the full panel."""
    import torch
    import torch.distributed as dist

    st = _STATE
    blocks = partition(R.ncols, st.world)
    c0, c1 = blocks[st.rank]
    widths = [b - a for a, b in blocks]
    wmax = max(widths)
    if c1 > c0:
        v2 = V2.cols(c0, c1).view if V2 is not None else empty_view
        be.check(be.lib.dre_adi_solve(be.h, mu.real, mu.imag, R.cols(c0, c1).view, V1.cols(c0, c1).view, v2))
    n = be.n
    trace = os.environ.get("DRE_DIST_TRACE") is not None
    lib_stream = _library_stream(be)
    if trace:
        import time

        t_host0 = time.perf_counter()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(lib_stream)
    with torch.cuda.stream(lib_stream):
        for V in (V1, V2):
            if V is None or wmax == 0:
                continue
            full = panel_tensor(be, V)
            # persistent exchange buffers: a fresh 160 MB tensor per call makes the caching allocator (which must
            # honour NCCL's record_stream) fall back to cudaMalloc -- a device synchronisation per ADI iteration
            key = (n, wmax, st.world, str(full.device))
            bufs = st.__dict__.setdefault("buffers", {})
            if key not in bufs:
                bufs.clear()
                bufs[key] = (torch.empty((n, wmax), dtype=torch.float64, device=full.device),
                             torch.empty((st.world * n, wmax), dtype=torch.float64, device=full.device))
            pad, out = bufs[key]
            pad[:, :c1 - c0].copy_(full[:, c0:c1])
            if c1 - c0 < wmax:
                pad[:, c1 - c0:].zero_()
            if trace:
                ev[1].record(lib_stream)
            dist.all_gather_into_tensor(out, pad, group=st.group)
            if trace:
                ev[2].record(lib_stream)
            for g, (b0, b1) in enumerate(blocks):
                if g != st.rank and b1 > b0:
                    full[:, b0:b1].copy_(out[g * n:(g + 1) * n, :b1 - b0])
            st.bytes_gathered += out.numel() * 8
            st.gathers += 1
    if trace:
        t_host1 = time.perf_counter()
        ev[2].synchronize()
        acc = st.__dict__.setdefault("trace", [0.0, 0.0, 0.0, 0])
        acc[0] += ev[0].elapsed_time(ev[1])   # solve (+ anything queued before) + pad copy
        acc[1] += ev[1].elapsed_time(ev[2])   # all-gather incl. waiting for the slowest rank
        acc[2] += 1e3 * (t_host1 - t_host0)   # host time to enqueue
        acc[3] += 1
        if acc[3] % 100 == 0 and st.rank == 0:
            print(f"[dre dist] per call: solve+pad {acc[0] / acc[3]:.3f} ms, all-gather(+wait) {acc[1] / acc[3]:.3f} ms, "
                  f"host enqueue {acc[2] / acc[3]:.3f} ms", flush=True)


def agree_scalar(x: float) -> float:
    """Rank 0's value of a control-flow scalar (all ranks compute it from identical data; this only
    guarantees that a discrepancy cannot desynchronise the collectives)."""
    if not active() or not _PARANOID:
        return x   # ranks are bit-identical by construction (tests/test_gpu_dist.py); checked at every compress!
    return float(agree_array(np.array([x], dtype=np.float64))[0])


def agree_array(a: np.ndarray) -> np.ndarray:
    """Broadcast rank 0's array (length first, then data); dtype float64 or complex128."""
    if not active():
        return a
    import torch
    import torch.distributed as dist

    st = _STATE
    dev = f"cuda:{st.device}" if st.device is not None else "cpu"
    a = np.ascontiguousarray(a)
    cplx = np.iscomplexobj(a)
    flat = a.astype(np.complex128).view(np.float64) if cplx else a.astype(np.float64)
    hdr = torch.tensor([flat.size, 1 if cplx else 0], dtype=torch.int64, device=dev)
    dist.broadcast(hdr, src=0, group=st.group)
    m, is_c = int(hdr[0].item()), bool(hdr[1].item())
    buf = torch.from_numpy(flat.copy()).to(dev) if st.rank == 0 else torch.empty(m, dtype=torch.float64, device=dev)
    if buf.numel() != m:
        buf = torch.empty(m, dtype=torch.float64, device=dev)
    dist.broadcast(buf, src=0, group=st.group)
    out = buf.cpu().numpy()
    return out.view(np.complex128) if is_c else out


def assert_same_int(v: int, what: str):
    """All ranks must hold the same integer (e.g. the rank after a compression); raises everywhere if not."""
    if not active():
        return
    import torch
    import torch.distributed as dist

    st = _STATE
    dev = f"cuda:{st.device}" if st.device is not None else "cpu"
    t = torch.tensor([v, -v], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=st.group)
    if int(t[0].item()) != -int(t[1].item()):
        raise RuntimeError(f"ranks disagree on {what}: max {int(t[0])} min {-int(t[1])}")
