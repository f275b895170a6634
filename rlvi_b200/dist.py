"""Sample-dimension sharding over the GPUs of one box: one process per GPU.

Every per-sample quantity (loss, e, pi, a row's contribution to the statistics) is independent, so rank r
owns the contiguous row block `shard_bounds(n, r, world)` of X, y, e, pi and nothing N-sized ever moves.
Only sums couple the shards:
  * the fixed point's three partial sums per pass are exchanged INSIDE the persistent kernel through
    NVLink peer windows (rlvi_dist_window_* + rlvi_fp_dist) and summed in rank order, so every rank gets
    the same bits and takes the same stop decision;
  * the d x d / d-vector statistics of the M-step are all-reduced once per step
    (`ShardGroup.all_reduce`): up to 8192 doubles through the library's own peer-window kernel
    (rlvi_stats_allreduce_f64 -- remote stores over NVLink, rank-ordered sum), larger ones with NCCL;
    gloo on CPU tensors in the unit tests.
The reference has no distributed code at all; this module is new.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as td

from . import _lib

__all__ = ["shard_bounds", "ShardGroup"]


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced row block of rank `rank`: sizes differ by at most one, even-sized first
    (so a 16-byte aligned FP64 vector stays 16-byte aligned at every shard start when n/world is even)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardGroup:
    """The ranks that share one sharded data set.  `create(device)` initialises torch.distributed from
    the torchrun environment if needed (127.0.0.1 rendezvous), allocates this rank's peer window on
    `device` (CUDA only) and maps the peers'."""

    def __init__(self, rank, world, device):
        self.rank, self.world, self.device = rank, world, device
        self.window = None
        self.table = None
        self.calls = 0
        self.stats_calls = 0
        self._ctx = None

    @classmethod
    def create(cls, device, backend=None):
        device = torch.device(device)
        if not td.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            kw = {}
            if device.type == "cuda":
                kw["device_id"] = device
            td.init_process_group(backend or ("nccl" if device.type == "cuda" else "gloo"), **kw)
        g = cls(td.get_rank(), td.get_world_size(), device)
        if device.type == "cuda" and g.world > 1:
            g._open_windows()
        return g

    # ---- NVLink peer windows -------------------------------------------------------------------
    def _open_windows(self):
        ctx = self._ctx = _lib.context(self.device.index)
        lib = ctx.lib
        win = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(lib.rlvi_dist_window_create(ctx.handle, self.world, C.byref(win), handle),
                   "rlvi_dist_window_create")
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(self.device)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        td.all_gather(allh, mine)
        blob = b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh)
        table = C.c_void_p()
        _lib.check(lib.rlvi_dist_window_open(ctx.handle, self.rank, self.world, win, blob, C.byref(table)),
                   "rlvi_dist_window_open")
        self.window, self.table = win, table
        td.barrier()

    def fp_dist(self, n_global: int):
        """A fresh rlvi_fp_dist for the NEXT fixed-point call (call_index advances in lock-step on all
        ranks: every rank must make the same sequence of sharded fixed-point calls)."""
        if self.window is None:
            raise RuntimeError("peer windows are only available on CUDA with world > 1")
        self.calls += 1
        return _lib.FpDist(self.rank, self.world, int(n_global), self.window, self.table, self.calls)

    def stats_dist(self):
        """A fresh rlvi_fp_dist for the NEXT statistics all-reduce (its own call_index sequence)."""
        if self.window is None:
            raise RuntimeError("peer windows are only available on CUDA with world > 1")
        self.stats_calls += 1
        return _lib.FpDist(self.rank, self.world, 0, self.window, self.table, self.stats_calls)

    # ---- statistics ----------------------------------------------------------------------------
    STATS_CAPACITY = 8192     # RLVI_DIST_STATS_CAPACITY

    def all_reduce(self, t: torch.Tensor):
        """In-place SUM over the ranks; every rank receives the same bits.  FP64 CUDA vectors of up to
        STATS_CAPACITY elements (the d*d + 2d + 2 statistics for d <= 88) go through the library's own
        peer-window kernel (rlvi_stats_allreduce_f64: remote stores over NVLink + rank-ordered sum, one
        launch); anything else through torch.distributed (NCCL on CUDA, gloo on CPU)."""
        if self.world == 1:
            return t
        if (self.window is not None and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
                and t.numel() <= self.STATS_CAPACITY):
            d = self.stats_dist()
            handle = torch.cuda.current_stream(self.device).cuda_stream
            ctx = _lib.context(self.device.index, handle)       # the launching stream's own context (scratch)
            _lib.check(ctx.lib.rlvi_stats_allreduce_f64(ctx.handle, C.c_void_p(t.data_ptr()), t.numel(), C.byref(d),
                                                        C.c_void_p(handle)), "rlvi_stats_allreduce_f64")
            return t
        td.all_reduce(t, op=td.ReduceOp.SUM)
        return t

    def close(self):
        if self.window is not None:
            torch.cuda.synchronize(self.device)
            td.barrier()
            self._ctx.lib.rlvi_dist_window_close(self._ctx.handle, self.rank, self.world, self.window, self.table)
            self.window = self.table = None
