"""Multi-GPU sharding of the posterior-sample loop.

Posterior samples are independent (posteriormodel.py:95-101), so rank g of G evaluates the
global sample ids [g*n/G, (g+1)*n/G) -- Philox is keyed by the GLOBAL id, so the union of the
shards is the same set of draws for every G -- and ONE all-reduce(SUM, fp32) over the [B,C]
moment buffers combines the predictive moments.  There is no other data-path collective.

On GPUs the all-reduce is this library's own one-kernel exchange over NVLink peer memory
(`dbx_comm_allreduce`, csrc/collective.cu: every rank's window is mapped by its peers through
CUDA IPC; sums are taken in rank order, so all ranks hold bit-identical results).  The process
group (NCCL) is used for the rendezvous only -- the all-gather of the 64-byte IPC handles -- and
as the transport when peer mapping is not possible (ranks on different nodes, no P2P): that
case is reported by `collective_kind()`.  CPU tensors (the gloo tests) go through
torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os

_comm = None          # PeerComm of the default process group, created on first use
_comm_failed = None   # reason the peer window could not be set up (then NCCL carries the exchange)
_last_kind = "none"


def _dist():
    import torch.distributed as dist
    return dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    try:
        dist = _dist()
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous, balanced split of sample ids 0..n-1: returns (offset, count)."""
    if world_size <= 1:
        return 0, n
    base, rem = divmod(n, world_size)
    cnt = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, cnt


class PeerComm:
    """The NVLink peer-memory communicator of this rank (one per process group)."""

    def __init__(self, rank, world_size, device_index, max_floats):
        from . import _lib
        self.lib = _lib.load()
        self._lib = _lib
        self.h = C.c_void_p()
        self.cap = int(max_floats)
        _lib.check(self.lib.dbx_comm_create(rank, world_size, device_index, self.cap, C.byref(self.h)))
        self.rank, self.world = rank, world_size

    def handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._lib.check(self.lib.dbx_comm_handle(self.h, buf))
        return buf.raw

    def connect(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        self._lib.check(self.lib.dbx_comm_connect(self.h, blob))

    def allreduce(self, a, b=None, stream=None):
        import torch
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream if stream is None else stream)
        self._lib.check(self.lib.dbx_comm_allreduce(self.h, C.c_void_p(a.data_ptr()), a.numel(),
                                                    C.c_void_p(b.data_ptr()) if b is not None else None,
                                                    b.numel() if b is not None else 0, st))

    def timed_out(self) -> bool:
        v = C.c_int32()
        self._lib.check(self.lib.dbx_comm_status(self.h, C.byref(v), None))
        return bool(v.value)

    def close(self):
        if self.h:
            self.lib.dbx_comm_destroy(self.h)
            self.h = C.c_void_p()


def _same_host(dist, ws):
    names = [None] * ws
    dist.all_gather_object(names, os.uname().nodename + ":" + str(os.stat("/proc/self/ns/ipc").st_ino if os.path.exists("/proc/self/ns/ipc") else 0))
    return len(set(names)) == 1


def peer_comm(max_floats=1 << 20):
    """The peer-memory communicator of the default group (None when it cannot be built: the reason is kept in
    `collective_kind()`).  Collective over the group: every rank must call it at the same point."""
    global _comm, _comm_failed
    if _comm is not None or _comm_failed is not None:
        return _comm
    import torch
    dist = _dist()
    rank, ws = world()
    if os.environ.get("DBX_NO_PEER_COMM") == "1":
        _comm_failed = "disabled by DBX_NO_PEER_COMM"
        return None
    ok, why, c = 1, "", None
    try:
        if not _same_host(dist, ws):
            raise RuntimeError("ranks span several hosts")
        c = PeerComm(rank, ws, torch.cuda.current_device(), max_floats)
        hs = [None] * ws
        dist.all_gather_object(hs, c.handle())
        c.connect(hs)
    except Exception as e:   # no P2P mapping between these ranks
        ok, why = 0, "%s: %s" % (type(e).__name__, e)
    flags = [None] * ws
    dist.all_gather_object(flags, (ok, why))
    if all(f[0] for f in flags):
        _comm = c
    else:
        if c is not None:
            c.close()
        _comm_failed = "; ".join(sorted({f[1] for f in flags if not f[0]}))[:300]
    return _comm


def collective_kind():
    """What carried the last all-reduce: 'nvlink_peer_oneshot' (this library's kernel), 'nccl' / 'gloo' (torch.distributed;
    for nccl the reason the peer window was not used follows), 'none'."""
    return _last_kind if not _comm_failed or _last_kind != "nccl" else "nccl (peer window unavailable: %s)" % _comm_failed


def allreduce_sum_(*tensors):
    """In-place SUM all-reduce of the moment buffers: ONE exchange for all of them."""
    global _last_kind
    rank, ws = world()
    ts = [t for t in tensors if t is not None]
    if ws <= 1 or not ts:
        return tensors
    import torch
    dist = _dist()
    if ts[0].is_cuda and len(ts) <= 2 and all(t.dtype == torch.float32 and t.is_contiguous() and t.numel() % 4 == 0 and
                                              t.data_ptr() % 16 == 0 for t in ts):
        c = peer_comm(max(1 << 20, sum(t.numel() for t in ts)))
        if c is not None and sum(t.numel() for t in ts) <= c.cap:
            c.allreduce(ts[0], ts[1] if len(ts) == 2 else None)
            _last_kind = "nvlink_peer_oneshot"
            return tensors
    _last_kind = "nccl" if ts[0].is_cuda else "gloo"
    if len(ts) == 1:
        dist.all_reduce(ts[0], op=dist.ReduceOp.SUM)
    else:
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        o = 0
        for t in ts:
            t.copy_(flat[o:o + t.numel()].view_as(t))
            o += t.numel()
    return tensors


def shutdown():
    """Unmap the peer windows (call before destroy_process_group)."""
    global _comm, _comm_failed, _last_kind
    if _comm is not None:
        _comm.close()
    _comm, _comm_failed, _last_kind = None, None, "none"


def sharded_upload_enabled(world_size=None) -> bool:
    """Is the input batch of a sharded predict uploaded in row shards + all-gather?  From 4 ranks on (measured on B200s with the
    12.8 MB batch of config 2: end-to-end step 1.24 -> 0.89 ms at 8 ranks, but 2.57 -> 2.69 ms at 2, where the all-gather costs more
    than half an upload saves); DBX_NO_SHARDED_UPLOAD=1 turns it off."""
    if os.environ.get("DBX_NO_SHARDED_UPLOAD", "0") not in ("", "0"):
        return False
    ws = world()[1] if world_size is None else world_size
    return ws >= int(os.environ.get("DBX_SHARDED_UPLOAD_MIN_RANKS", "4"))


def sharded_upload(x_rows, upload, device):
    """The SAME host batch ``x_rows`` ([B,K] fp32, every rank passes it: the sharded predict evaluates all samples on one input)
    as a [B,K] device tensor on every rank: rank g uploads rows [g*ceil(B/G), ...) over its own PCIe link (``upload(host_slice)
    -> device tensor``) and ONE all-gather over NVLink / NVSwitch completes the batch -- instead of G uploads of the whole batch
    (at 8 ranks the 12.8 MB batch of config 2 cost each rank 0.5 ms per call, the whole of its compute share).  Returns None when
    there is nothing to shard."""
    import torch
    rank, ws = world()
    B, K = int(x_rows.shape[0]), int(x_rows.shape[1])
    if ws <= 1 or B < ws:
        return None
    per = (B + ws - 1) // ws
    lo = min(rank * per, B)
    hi = min(lo + per, B)
    mine = upload(x_rows[lo:hi]) if hi > lo else torch.empty((0, K), dtype=torch.float32, device=device)
    if hi - lo < per:                                   # ragged last shard: the gather needs equal pieces
        pad = torch.zeros((per, K), dtype=torch.float32, device=device)
        pad[:hi - lo] = mine
        mine = pad
    full = torch.empty((per * ws, K), dtype=torch.float32, device=device)
    _dist().all_gather_into_tensor(full, mine.contiguous())
    return full[:B]

