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


# =============================================================================================
# Pipeline mode (default for world_size > 1): the two lanes of the ADI loop on two GPUs
# =============================================================================================
# What a time step costs on ONE GPU (n = 79 841, profiles/r02_results.md): the latency-bound ADI iteration chain
# (factor -> sweeps -> SMW -> SpMM -> residual norm, ~2.6 ms x 100) and the column compression compress!(X) of
# adi.jl:143-147 (rank-revealing Gram-Schmidt rounds + fat DMMA passes, ~35 ms x 12) are about the same size, and
# nothing in the iteration reads X: a step only appends the term (-2 mu alpha) V T V' (adi.jl:170-176).  Sharding
# the RHS columns of the solves over the ranks ("columns" mode above, DRE_DIST_MODE=columns) therefore divides the
# smaller half of the step and pays an all-gather per iteration; it was measured to scale negatively.  Pipeline mode
# gives each lane its own GPU instead:
#   rank 0  runs the reference's drivers unchanged and streams every increment (V panel + core) to rank 1 right
#           after the step that produced it (NCCL send on the library's stream, asynchronous);
#   rank 1  holds X: appends the terms, runs compress!(X) whenever rank 0's loop reaches a compression point
#           (adi.jl:112-114) and returns the compressed factor when X is read (end of the ADI solve);
#   ranks >= 2 have no lane of this path to run and idle until the job ends.
# Control messages travel over a gloo group (CPU tensors), panels over NCCL.  Same arithmetic on the same operands
# in the same order as the single-GPU path: results are identical.
_PIPE = None
# NCCL aborts when two host threads of one process launch on the same device at the same time ("host threads racing to
# launch NCCL on same device", seen with two lanes: the receiver of rank 0's terms and the receiver of the other lane's
# hand-over).  Every NCCL enqueue of a lane goes through this lock; the stream synchronisations stay outside it.
import threading as _threading

_NCCL_LOCK = _threading.Lock()


def _nccl_lock():
    """The lock for NCCL enqueues (asynchronous: held for microseconds); nothing for gloo, whose recv blocks."""
    import contextlib

    return _NCCL_LOCK if _PIPE is not None and _PIPE.data_backend != "gloo" else contextlib.nullcontext()
CMD_STOP, CMD_BEGIN, CMD_TERM, CMD_COMPRESS, CMD_FETCH, CMD_PREV = 0, 1, 2, 3, 4, 5
_HDR = 8


class _Pipe:
    def __init__(self, rank, world, device, ctl, data_backend):
        self.rank, self.world, self.device, self.ctl, self.data_backend = rank, world, device, ctl, data_backend
        self.token = 0            # identifies the compressed factor the holding lane currently has as its first term
        self.sends = []           # outstanding isend works (and the panels they read)
        # DRE_PIPE_LANES=2 (needs >= 3 ranks): two compression lanes take the compression points in turn, see serve()
        self.nlanes = max(1, min(world - 1, int(os.environ.get("DRE_PIPE_LANES", "1"))))
        self.lanes = list(range(1, 1 + self.nlanes))
        self.cur = 1              # rank 0: the lane the next terms go to (= the lane that holds X between solves)
        self.stats = {"terms_sent": 0, "bytes_sent": 0, "compress_cmds": 0, "fetches": 0, "fetch_wait_s": 0.0}


def enable_pipeline(device=None):
    """All ranks call this once after torch.distributed is initialised (NCCL for GPUs, gloo for the CPU tests)."""
    global _PIPE
    import datetime

    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("dre_b200.dist.enable_pipeline: torch.distributed is not initialised")
    rank, world = dist.get_rank(), dist.get_world_size()
    data_backend = dist.get_backend()
    ctl = None
    if data_backend != "gloo":
        ctl = dist.new_group(ranks=list(range(world)), backend="gloo", timeout=datetime.timedelta(seconds=900))
    _PIPE = _Pipe(rank, world, device, ctl, data_backend)
    return _PIPE


def pipeline_active() -> bool:
    return _PIPE is not None and _PIPE.world > 1 and _PIPE.rank == 0


def pipe_state():
    return _PIPE


def _dev_str():
    return "cpu" if _PIPE.data_backend == "gloo" else f"cuda:{_PIPE.device if _PIPE.device is not None else 0}"


def _send_hdr(dst, vals):
    import torch
    import torch.distributed as dist

    h = torch.zeros(_HDR, dtype=torch.float64)
    h[:len(vals)] = torch.tensor([float(v) for v in vals], dtype=torch.float64)
    dist.send(h, dst=dst, group=_PIPE.ctl)


def _recv_hdr(src):
    import torch
    import torch.distributed as dist

    h = torch.zeros(_HDR, dtype=torch.float64)
    dist.recv(h, src=src, group=_PIPE.ctl)
    return h.tolist()


def _panel_io(be, M):
    """(tensor aliasing the panel, context manager that makes torch's current stream the library's stream)."""
    import contextlib

    import torch

    if _PIPE.data_backend == "gloo":        # CPU tests: the emulator's "device" memory is host memory
        ptr, ld = C.c_void_p(), C.c_int64()
        be.check(be.lib.dre_mat_devptr(be.h, M.view, C.byref(ptr), C.byref(ld)))
        be.check(be.lib.dre_sync(be.h))
        buf = (C.c_double * (be.n * ld.value)).from_address(ptr.value)
        full = torch.from_numpy(np.frombuffer(buf, dtype=np.float64).reshape(be.n, ld.value))
        return full[:, :M.ncols], contextlib.nullcontext()
    ptr, ld = C.c_void_p(), C.c_int64()
    be.check(be.lib.dre_mat_devptr(be.h, M.view, C.byref(ptr), C.byref(ld)))
    dev = _PIPE.device if _PIPE.device is not None else 0
    t = torch.as_tensor(_CudaArray(ptr.value, be.n, M.ncols, ld.value), device=f"cuda:{dev}")
    sp = C.c_void_p()
    be.check(be.lib.dre_get_stream(be.h, C.byref(sp)))
    return t, torch.cuda.stream(torch.cuda.ExternalStream(sp.value, device=f"cuda:{dev}"))


def _send_panel(be, M, dst):
    """Asynchronous send of an n x k panel, ordered behind the kernels that wrote it; the panel is kept alive
    until the send has completed (reaped at the next compression point / fetch)."""
    import torch
    import torch.distributed as dist

    t, on_stream = _panel_io(be, M)
    with on_stream:
        if _PIPE.data_backend == "gloo":
            t = t.contiguous()
        elif not t.is_contiguous():
            # a column view of a wider panel / a padded leading dimension: NCCL wants one dense buffer.  Staged through
            # a small ring of persistent buffers (a fresh 150 MB torch tensor per ADI step would push the caching
            # allocator into cudaMalloc / cudaFree, i.e. device synchronisations on the ADI chain).
            ring = _PIPE.__dict__.setdefault("ring", {})
            slots = ring.setdefault(tuple(t.shape), [])
            buf = None
            for sl in slots:
                if sl[1] is None or sl[1].is_completed():
                    buf = sl
                    break
            if buf is None:
                buf = [torch.empty(t.shape, dtype=t.dtype, device=t.device), None]
                slots.append(buf)
            buf[0].copy_(t)
            t = buf[0]
            with _nccl_lock():
                w = dist.isend(t, dst=dst)
            buf[1] = w
            _PIPE.sends.append((w, M, t))
            _PIPE.stats["bytes_sent"] += t.numel() * 8
            return
        with _nccl_lock():
            w = dist.isend(t, dst=dst)
    _PIPE.sends.append((w, M, t))
    _PIPE.stats["bytes_sent"] += t.numel() * 8


def _reap_sends(block=False):
    keep = []
    for w, M, t in _PIPE.sends:
        if block:
            w.wait()
        elif not w.is_completed():
            keep.append((w, M, t))
    _PIPE.sends = keep


def _recv_panel(be, M, src):
    import torch
    import torch.distributed as dist

    t, on_stream = _panel_io(be, M)
    with on_stream:
        if t.is_contiguous():
            dist.recv(t, src=src)
        else:
            tmp = torch.empty(t.shape, dtype=t.dtype, device=t.device)
            dist.recv(tmp, src=src)
            t.copy_(tmp)
        if _PIPE.data_backend != "gloo":
            torch.cuda.current_stream().synchronize()


def _send_small(arr, dst):
    """Small float64 array (cores, eigenvalues) over the control group."""
    import torch
    import torch.distributed as dist

    a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64)).ravel()
    if a.size:
        dist.send(torch.from_numpy(a.copy()), dst=dst, group=_PIPE.ctl)


def _recv_small(count, src):
    import torch
    import torch.distributed as dist

    out = torch.zeros(int(count), dtype=torch.float64)
    if count:
        dist.recv(out, src=src, group=_PIPE.ctl)
    return out.numpy()


# ---- rank 0 side -------------------------------------------------------------------------------------------------
def _scale_of(X0):
    """sqrt of the largest |alpha d_j| over the diagonal-core terms of X0 (0 if unknown): the column scale compress!
    measures its drop threshold against; a lane that starts a job without X needs it as a hint."""
    sc = 0.0
    for a, L, D in zip(X0.alphas, X0.Ls, X0.Ds):
        D = np.asarray(D)
        if L.ncols and D.ndim == 2 and D.shape[0] and np.count_nonzero(D - np.diag(np.diag(D))) == 0:
            sc = max(sc, float(np.sqrt(np.max(np.abs(a * np.diag(D))))))
    return sc


def pipe_begin(be, X0):
    """Start of an ADI solve (adi.jl:29-69): the holding lane's X becomes the initial guess.  If the initial guess is
    the very factor that lane returned last (Ros1/Ros2 pass the previous X, lowrank_ros1.jl:48-49) nothing is
    transferred."""
    p = _PIPE
    same = False
    if len(X0.Ls) == 1 and p.token != 0 and getattr(X0.Ls[0], "_pipe_token", None) == p.token and X0.alphas[0] == 1.0:
        D0 = np.asarray(X0.Ds[0])
        lam = getattr(X0.Ls[0], "_pipe_lam", None)
        same = (lam is not None and D0.shape == (lam.size, lam.size) and np.array_equal(np.diag(D0), lam)
                and np.count_nonzero(D0 - np.diag(np.diag(D0))) == 0)
    has_base = same or any(L.ncols for L in X0.Ls)
    scale = _scale_of(X0)
    for lane in p.lanes:
        mine = lane == p.cur
        _send_hdr(lane, [CMD_BEGIN, be.n, 1.0 if (same and mine) else 0.0, 1.0 if (has_base and mine) else 0.0, scale])
    if not same:
        p.token = 0
        for a, L, D in zip(X0.alphas, X0.Ls, X0.Ds):
            if L.ncols:
                pipe_send_term(be, a, L, D)


def pipe_send_term(be, alpha, L, D):
    dst = _PIPE.cur
    D = np.asfortranarray(np.asarray(D, dtype=np.float64))
    diag = np.count_nonzero(D - np.diag(np.diag(D))) == 0
    _send_hdr(dst, [CMD_TERM, L.ncols, float(alpha), 1.0 if diag else 0.0])
    _send_small(np.diag(D) if diag else D, dst)
    _send_panel(be, L, dst)
    _PIPE.stats["terms_sent"] += 1
    _reap_sends()


def pipe_compress():
    """Compression point (adi.jl:143-147): the current lane compresses what it holds and -- with two lanes -- hands
    the result to the other lane, which receives the following terms."""
    p = _PIPE
    nxt = p.lanes[(p.lanes.index(p.cur) + 1) % p.nlanes]
    _send_hdr(p.cur, [CMD_COMPRESS, nxt])
    p.cur = nxt
    p.stats["compress_cmds"] += 1


def pipe_fetch(be, make_panel):
    """compress!(X) on the current lane (if more than one term is pending) and its result: (DeviceMatrix, eigenvalues)."""
    import time

    p = _PIPE
    src = p.cur
    p.fetch_seq = getattr(p, "fetch_seq", 0) + 1
    _send_hdr(src, [CMD_FETCH, p.fetch_seq])   # the token that will identify the returned factor (unique over lanes)
    t0 = time.perf_counter()
    h = _recv_hdr(src)
    k2, token = int(h[0]), int(h[1])
    lam = _recv_small(k2, src)
    L = make_panel(k2)
    if k2:
        _recv_panel(be, L, src)
    p.stats["fetch_wait_s"] += time.perf_counter() - t0
    p.stats["fetches"] += 1
    _reap_sends(block=True)
    p.token = token
    L._pipe_token = token
    L._pipe_lam = np.array(lam, copy=True)
    return L, lam


def pipe_stop():
    """Rank 0 releases the other ranks (end of the job)."""
    if _PIPE is None or _PIPE.world < 2 or _PIPE.rank != 0:
        return
    _reap_sends(block=True)
    for r in range(1, _PIPE.world):
        _send_hdr(r, [CMD_STOP])


# ---- lanes (rank 1, or ranks 1 and 2) and the idle ranks ------------------------------------------------------------
class _TensorPanel:
    """Panel of the lane's context that aliases a torch tensor (n x ld, float64) filled by the receiver thread
    (dre_mat_wrap).  Synchronises the context before the tensor goes back to torch's allocator."""

    def __init__(self, be, tensor, cols):
        self.be, self.cols, self.gen, self.tensor = be, cols, be.generation, tensor
        self.version = 0
        pid = C.c_int32(-1)
        be.check(be.lib.dre_mat_wrap(be.h, C.c_void_p(tensor.data_ptr()), int(tensor.stride(0)), cols, C.byref(pid)))
        self.id = pid.value

    def __del__(self):
        try:
            if self.be.ctx.h and self.gen == self.be.generation:
                self.be.lib.dre_sync(self.be.h)
                self.be.lib.dre_mat_free(self.be.h, self.id)
        except Exception:
            pass


def _recv_panel_tensor(n, k, src):
    """Receiver thread of rank 1: an n x k panel into a fresh torch tensor with an even leading dimension (the
    library wants 16-byte rows); complete when this returns."""
    import torch
    import torch.distributed as dist

    dev = _dev_str()
    kp = k + (k & 1)
    # every buffer has one of a few sizes (columns rounded up to 64): torch's caching allocator then recycles the
    # blocks of compressed-away terms instead of calling cudaMalloc (a device synchronisation in the middle of the
    # running compress!) for each new column count
    cap = n * ((kp + 63) // 64 * 64)
    flat = torch.empty(cap, dtype=torch.float64, device=dev)
    if kp == k:
        t = flat[: n * k].view(n, k)
        with _nccl_lock():
            dist.recv(t, src=src)
    else:
        tmp = torch.empty(cap, dtype=torch.float64, device=dev)[: n * k].view(n, k)
        with _nccl_lock():
            dist.recv(tmp, src=src)
        t = flat[: n * kp].view(n, kp)
        t[:, :k].copy_(tmp)
        t[:, k:].zero_()
    if _PIPE.data_backend != "gloo":
        torch.cuda.current_stream().synchronize()
    return t


def serve(api):
    """Rank >= 1: a compression lane or an idle wait for the end of the job.

    A lane runs several host threads.  The RECEIVERS (one per peer that can send to it: rank 0, and the other lane
    when there are two) post every receive as soon as the peer announces it (control headers and cores over gloo --
    whose send is a rendezvous: it completes only once the peer has posted the matching receive -- and panels over
    NCCL into fresh torch tensors) and queue the messages in order; the WORKER (this thread) appends terms and runs
    compress!.  With a single thread rank 0's sends waited for the running compress! and the two GPUs ran one after
    the other (profiles/r02_results.md: 1.26 steps/s on two GPUs, the same as on one).

    TWO LANES (DRE_PIPE_LANES=2, >= 3 ranks).  compress! number k+1 needs X_k, the result of number k, but most of
    its work does not: the Gram-Schmidt of its ten increments among themselves.  So the lanes take the compression
    points in turn; rank 0 sends the increments of k+1 to the lane that did not get those of k; at its compression
    point that lane orthogonalises them first (dre_compress_begin / _add), then adds X_k -- handed over by the other
    lane when its compress! ends -- as the LAST term (one more _add), and finishes.  Only that tail (~10 ms) is on
    the lane-to-lane dependency chain, each lane has twice the time per compress!.  Any term order gives the same
    X up to round-off (the orthonormal basis spans the same space); the drop threshold of the basis is relative to
    the largest column met, which the increments-first order would only learn at the end, so the lane passes the
    scale of the last X it knows as a hint (dre_compress_scale_hint)."""
    p = _PIPE
    if p.rank not in p.lanes:
        while int(_recv_hdr(0)[0]) != CMD_STOP:
            pass
        return {"role": "idle"}
    import queue
    import threading
    import time

    me = p.rank
    others = [r for r in p.lanes if r != me]
    be = None
    terms = []        # (alpha, DeviceMatrix, core) in arrival order
    base_ok = False   # the terms at hand include X (the initial guess, or this lane's / the other lane's last result)
    scale_hint = 0.0
    served = {"role": "compress", "terms": 0, "compressions": 0, "fetches": 0, "busy_s": 0.0, "idle_s": 0.0,
              "prev_wait_s": 0.0, "handovers": 0}
    inbox = queue.Queue()      # commands of rank 0, in order
    prevbox = queue.Queue()    # X_k handed over by the other lane

    def cuda_thread_setup():
        if p.data_backend != "gloo":
            import torch

            torch.cuda.set_device(p.device if p.device is not None else 0)
            torch.cuda.set_stream(torch.cuda.Stream())

    def receiver():
        try:
            cuda_thread_setup()
            n = 0
            while True:
                h = _recv_hdr(0)
                cmd = int(h[0])
                if cmd == CMD_BEGIN:
                    n = int(h[1])
                if cmd == CMD_TERM:
                    k, diag = int(h[1]), h[3] != 0.0
                    core = _recv_small(k if diag else k * k, 0)
                    inbox.put((cmd, h, core, _recv_panel_tensor(n, k, 0)))
                else:
                    inbox.put((cmd, h, None, None))
                if cmd == CMD_STOP:
                    return
        except BaseException as e:  # noqa: BLE001  (the worker re-raises it)
            inbox.put((-1, e, None, None))

    def prev_receiver(src):
        try:
            cuda_thread_setup()
            while True:
                h = _recv_hdr(src)
                if int(h[0]) == CMD_STOP:
                    return
                k, n = int(h[1]), int(h[2])
                lam = _recv_small(k, src)
                prevbox.put((lam, _recv_panel_tensor(n, k, src) if k else None, k))
        except BaseException as e:  # noqa: BLE001
            prevbox.put((e, None, -1))

    threads = [threading.Thread(target=receiver, name="dre-pipe-receiver", daemon=True)]
    threads += [threading.Thread(target=prev_receiver, args=(r,), name=f"dre-pipe-prev-{r}", daemon=True) for r in others]
    for th in threads:
        th.start()

    # Streaming compress! (DRE_PIPE_STREAM=1, opt-in): every term is handed to the open job (api.CompressStream =
    # dre_compress_begin / _add / _finish) as soon as it has arrived, so the lane orthogonalises increment i while
    # rank 0 computes i+1 and only the core / eigen / L <- QV tail is left when the compression point comes.
    # Measured on two B200s (profiles/r02_results.md, run r02w): 1.69 steps/s against 1.78 with one
    # dre_ldlt_compress call per compression point -- one-term jobs lose the look-ahead overlap of the chunk pipeline
    # (lane busy 47 instead of 43 ms per compress!), the lane stays the longer of the two lanes either way, and the
    # different projection order changes round-off, i.e. the free-running shift sequence (the 2-GPU run is no longer
    # bit-identical to the 1-GPU run).  It becomes the right default once compress! is shorter than the ten
    # iterations it covers.
    streaming = os.environ.get("DRE_PIPE_STREAM", "0") not in ("", "0")
    job = None       # open CompressStream; job_n = how many of `terms` it holds
    job_n = 0

    def open_job(live):
        widest = max(L.ncols for _, L, _ in live)
        j = api.CompressStream(be, sum(L.ncols for _, L, _ in live) + 16 * max(widest, 64))
        if not base_ok and scale_hint > 0.0:
            j.scale_hint(scale_hint)
        return j

    def feed():
        """Hand the terms that are not in the job yet to it (opening one when at least two terms exist)."""
        nonlocal job, job_n
        if not streaming or be is None:
            return
        live = [(a, L, D) for a, L, D in terms if L.ncols]
        if job is None:
            if len(live) < 2:
                return
            # room for the terms at hand plus a dozen increments of the widest kind seen so far; if a solve ever
            # needs more, the job is dropped and the compression point falls back to the one-call path
            job = open_job(live)
            job_n = 0
        new = live[job_n:]
        if not new:
            return
        if not job.room_for(sum(L.ncols for _, L, _ in new)):
            job = None
            job_n = 0
            return
        t0 = time.perf_counter()
        job.add(new)
        served["busy_s"] += time.perf_counter() - t0
        job_n = len(live)

    def take_prev():
        """X_k from the other lane (blocks until its compress! has ended); appended as the last term."""
        nonlocal base_ok, scale_hint
        t0 = time.perf_counter()
        lam, tensor, k = prevbox.get()
        served["prev_wait_s"] += time.perf_counter() - t0
        if k < 0:
            raise lam
        if k:
            L = api.DeviceMatrix(_TensorPanel(be, tensor, k), 0, k)
            api._mark_orthonormal(L)      # (the outer factor of a compress!)
            terms.append((1.0, L, np.asfortranarray(np.diag(lam))))
            scale_hint = max(scale_hint, float(np.sqrt(np.max(np.abs(lam)))))
        base_ok = True

    def compress_now():
        nonlocal terms, job, job_n, scale_hint
        if not base_ok:
            # increments first, X last: everything that does not need X runs before the hand-over is awaited
            live = [(a, L, np.asfortranarray(D)) for a, L, D in terms if L.ncols]
            if live:
                t0 = time.perf_counter()
                if job is None:
                    job = open_job(live)
                    job_n = 0
                if job.room_for(sum(L.ncols for _, L, _ in live[job_n:])):
                    job.add(live[job_n:])
                    job_n = len(live)
                else:
                    job, job_n = None, 0
                served["busy_s"] += time.perf_counter() - t0
            take_prev()
        live = [(a, L, np.asfortranarray(D)) for a, L, D in terms if L.ncols]
        if len(live) == 1 and api._is_orthonormal(live[0][1]):
            job, job_n = None, 0
            return
        if not live:
            job, job_n = None, 0
            return
        feed()
        t0 = time.perf_counter()
        if job is not None and job.room_for(sum(L.ncols for _, L, _ in live[job_n:])):
            if job_n < len(live):
                job.add(live[job_n:])
            Lnew, lam = job.finish()
            served["streamed"] = served.get("streamed", 0) + 1
        else:
            Lnew, lam = api._compress_call(be, live)
        job, job_n = None, 0
        be.ctx.sync()
        served["busy_s"] += time.perf_counter() - t0
        served["compressions"] += 1
        terms = [(1.0, Lnew, np.asfortranarray(np.diag(lam)))]
        if lam.size:
            scale_hint = max(scale_hint, float(np.sqrt(np.max(np.abs(lam)))))

    while True:
        t0 = time.perf_counter()
        cmd, h, core, tensor = inbox.get()
        served["idle_s"] += time.perf_counter() - t0
        if cmd == -1:
            raise h
        if cmd == CMD_STOP:
            for r in others:           # releases the hand-over receiver of the other lane
                _send_hdr(r, [CMD_STOP])
            break
        if cmd == CMD_BEGIN:
            n = int(h[1])
            if be is None or be.n != n:
                terms = []
                job, job_n = None, 0
                be = api._pipe_lane_backend(n, p.device)
            if h[2] == 0.0:
                terms = []
                job, job_n = None, 0
            base_ok = h[3] != 0.0
            scale_hint = float(h[4])
        elif cmd == CMD_TERM:
            k, alpha, diag = int(h[1]), h[2], h[3] != 0.0
            D = np.diag(core) if diag else core.reshape(k, k, order="F")
            L = api.DeviceMatrix(_TensorPanel(be, tensor, k), 0, k)
            terms.append((alpha, L, np.asfortranarray(D)))
            served["terms"] += 1
            feed()
        elif cmd == CMD_COMPRESS:
            nxt = int(h[1]) if len(h) > 1 and h[1] else me
            compress_now()
            if nxt != me:
                # hand X_k over to the lane that is collecting the increments of the next compression point
                if terms:
                    _, L, D = terms[0]
                    lam = np.diag(D).copy()
                else:
                    L, lam = None, np.zeros(0)
                k2 = 0 if L is None else L.ncols
                _send_hdr(nxt, [CMD_PREV, k2, be.n])
                _send_small(lam, nxt)
                if k2:
                    _send_panel(be, L, nxt)
                    _reap_sends(block=True)
                terms = []
                base_ok = False
                served["handovers"] += 1
            else:
                base_ok = True
        elif cmd == CMD_FETCH:
            compress_now()
            token = int(h[1]) if len(h) > 1 and h[1] else p.token + 1
            p.token = token
            if terms:
                _, L, D = terms[0]
                lam = np.diag(D).copy()
            else:
                L, lam = None, np.zeros(0)
            k2 = 0 if L is None else L.ncols
            _send_hdr(0, [k2, token])
            _send_small(lam, 0)
            if k2:
                _send_panel(be, L, 0)
                _reap_sends(block=True)
            served["fetches"] += 1
    for th in threads:
        th.join(timeout=5.0)
    return served
