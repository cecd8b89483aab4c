"""Multi-GPU layer: one process per GPU, torch.distributed for the plumbing.

SURVEY.md 8e: reads are independent units (mash.go:68-104 touches only its receiver), so
rank r of P sketches reads [lo_r, hi_r) with NO communication.  Distance needs exactly one
exchange step: an all-gather of the finished (compact) sketches over NCCL/NVLink, after which
rank r computes row block r of the pair matrix against the gathered set (receiver = row).

The communication pattern is independent of who computes: `sharded_sketch_distance` takes
the two compute callables, and defaults to the CUDA entry points of libpolyb200.so.  (The
CPU-tier test drives the same function over gloo with a CPU stand-in injected by the test
itself, to check the partition / gather / row-block bookkeeping without a GPU.)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition of n items; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass
class ShardPlan:
    n_total: int
    rank: int
    world: int

    @property
    def lo(self) -> int:
        return shard_range(self.n_total, self.rank, self.world)[0]

    @property
    def hi(self) -> int:
        return shard_range(self.n_total, self.rank, self.world)[1]

    @property
    def counts(self):
        return [shard_range(self.n_total, r, self.world)[1] - shard_range(self.n_total, r, self.world)[0] for r in range(self.world)]


def all_gather_rows(local: torch.Tensor, plan: ShardPlan, group=None) -> torch.Tensor:
    """All-gather row-sharded [n_r, w] tensors into [n_total, w] (rank order == row order).
    Equal shards use one all_gather_into_tensor (ncclAllGather); ragged shards pad to the
    largest shard and trim."""
    counts = plan.counts
    w = local.shape[1]
    if plan.world == 1:
        return local
    view = local.view(torch.int32) if local.dtype == torch.uint32 else local
    if len(set(counts)) == 1:
        out = torch.empty((plan.n_total, w), dtype=view.dtype, device=view.device)
        dist.all_gather_into_tensor(out, view.contiguous(), group=group)
    else:
        m = max(counts)
        padded = torch.zeros((m, w), dtype=view.dtype, device=view.device)
        padded[: view.shape[0]] = view
        buf = torch.empty((plan.world * m, w), dtype=view.dtype, device=view.device)
        dist.all_gather_into_tensor(buf, padded, group=group)
        out = torch.cat([buf[r * m: r * m + counts[r]] for r in range(plan.world)], dim=0)
    return out.view(local.dtype) if local.dtype == torch.uint32 else out


# ---- default (CUDA) compute callables ---------------------------------------------------
def cuda_sketch_uniform(d_bases: torch.Tensor, n_reads: int, read_len: int, k: int, s: int) -> torch.Tensor:
    """Device-resident reads -> device-resident compact sketches [n, min(max(L-k,0), s)] (int32 view)."""
    from . import _lib

    cnt = min(max(read_len - k, 0), s)
    out = torch.empty((n_reads, max(cnt, 1)), dtype=torch.int32, device=d_bases.device)
    _lib.check(_lib.lib().pg_mash_sketch_uniform_dev(d_bases.data_ptr(), n_reads, read_len, k, s, 0, out.data_ptr(), max(cnt, 1),
                                                      None, torch.cuda.current_stream().cuda_stream))
    return out[:, :cnt] if cnt else out[:, :0]


def cuda_distance_block(d_sketches: torch.Tensor, row_begin: int, row_end: int) -> torch.Tensor:
    """Rows [row_begin,row_end) x all columns matching counts (uint32 as int32), literal reference semantics."""
    from . import _lib

    n, s = d_sketches.shape
    same = torch.empty((row_end - row_begin, n), dtype=torch.int32, device=d_sketches.device)
    _lib.check(_lib.lib().pg_mash_distance_block_dev(d_sketches.contiguous().data_ptr(), n, s, row_begin, row_end, same.data_ptr(), None,
                                                      torch.cuda.current_stream().cuda_stream))
    return same


def sharded_sketch_distance(local_reads: torch.Tensor, plan: ShardPlan, read_len: int, k: int, s: int,
                            sketch_fn: Optional[Callable] = None, distance_fn: Optional[Callable] = None, group=None,
                            pad_to_sketch_size: bool = True):
    """Sketch the local shard, all-gather the sketches, compute this rank's row block.

    Returns (local_sketches [n_r, cnt], gathered [n_total, s or cnt], same [n_r, n_total]).
    With pad_to_sketch_size the gathered rows are the full Go arrays (zero tail of a fresh
    Mash materialised AFTER the gather, so only informative words cross NVLink)."""
    sketch_fn = sketch_fn or cuda_sketch_uniform
    distance_fn = distance_fn or cuda_distance_block
    n_local = plan.hi - plan.lo
    local = sketch_fn(local_reads, n_local, read_len, k, s)
    gathered = all_gather_rows(local, plan, group)
    if pad_to_sketch_size and gathered.shape[1] < s:
        full = torch.zeros((plan.n_total, s), dtype=gathered.dtype, device=gathered.device)
        full[:, : gathered.shape[1]] = gathered
        gathered = full
    same = distance_fn(gathered, plan.lo, plan.hi)
    return local, gathered, same


# ---- fused sketch + all-gather over NVLink peer memory ------------------------------------
class GatheredBuffer:
    """A cudaMalloc'ed [n_total, cnt] uint32 buffer on this rank, exported to / mapped from every
    peer process through CUDA IPC handles exchanged over torch.distributed.  Every rank sizes its
    buffer from plan.n_total and rank r owns rows [plan.lo, plan.hi): shards may differ in size."""

    def __init__(self, plan: ShardPlan, cnt: int, group=None):
        import ctypes as C

        from . import _lib

        self._lib, self.plan, self.cnt = _lib, plan, cnt
        L = _lib.lib()
        self.nbytes = max(plan.n_total * max(cnt, 1) * 4, 16)
        p = C.c_void_p()
        _lib.check(L.pg_dev_alloc(C.byref(p), self.nbytes))
        self.ptr = p.value
        handle = (C.c_uint8 * 64)()
        _lib.check(L.pg_ipc_export(self.ptr, handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        allh = torch.empty((plan.world, 64), dtype=torch.uint8, device=mine.device)
        if plan.world > 1:
            dist.all_gather_into_tensor(allh, mine, group=group)
        else:
            allh[0] = mine
        allh = allh.cpu().numpy()
        self.peer_ptrs = []
        for r in range(plan.world):
            if r == plan.rank:
                self.peer_ptrs.append(self.ptr)
            else:
                q = C.c_void_p()
                h = (C.c_uint8 * 64)(*allh[r].tolist())
                _lib.check(L.pg_ipc_import(h, C.byref(q)))
                self.peer_ptrs.append(q.value)
        self._arr = (C.c_void_p * plan.world)(*self.peer_ptrs)

    def as_tensor(self) -> torch.Tensor:
        """Zero-copy int32 view [n_total, cnt] of this rank's gathered buffer."""
        class _Iface:
            pass

        h = _Iface()
        h.__cuda_array_interface__ = {"shape": (self.plan.n_total, max(self.cnt, 1)), "typestr": "<i4", "data": (self.ptr, False), "version": 2}
        return torch.as_tensor(h, device=torch.device("cuda", torch.cuda.current_device()))[:, : self.cnt]

    def to_numpy(self) -> np.ndarray:
        out = np.empty((self.plan.n_total, max(self.cnt, 1)), dtype=np.uint32)
        self._lib.check(self._lib.lib().pg_memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes, None))
        self._lib.check(self._lib.lib().pg_stream_sync(None))
        return out[:, : self.cnt]

    def close(self, group=None):
        L = self._lib.lib()
        torch.cuda.synchronize()
        if self.plan.world > 1:
            dist.barrier(group=group)  # nobody may still be storing into a buffer that is about to go
        for r, q in enumerate(self.peer_ptrs):
            if r != self.plan.rank:
                L.pg_ipc_close(q)
        L.pg_dev_free(self.ptr)


def fused_sketch_gather(d_bases: torch.Tensor, read_len: int, k: int, s: int, gathered: GatheredBuffer, group=None, sync: bool = True):
    """ONE kernel sketches this rank's reads [plan.lo, plan.hi) and stores every finished tile / row
    into rows [plan.lo, plan.hi) of the gathered buffer of every rank (TMA bulk stores to peer-mapped
    addresses in the fill regime): the all-gather overlaps the hashing tile by tile.  With sync it
    returns after a stream sync + barrier, i.e. when this rank's gathered buffer is complete."""
    from . import _lib

    plan = gathered.plan
    _lib.check(_lib.lib().pg_mash_sketch_uniform_scatter_dev(d_bases.data_ptr(), plan.hi - plan.lo, read_len, k, s, gathered._arr, plan.world,
                                                              plan.rank, plan.lo, torch.cuda.current_stream().cuda_stream))
    if sync:
        torch.cuda.synchronize()
        if plan.world > 1:
            dist.barrier(group=group)
