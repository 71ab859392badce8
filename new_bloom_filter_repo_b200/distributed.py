"""
Multi-GPU plumbing: one process per GPU.  Inter-frame pairs are independent units
(BloomFilterCompressor.compress keeps no state between frames, ivc:198-266), so ranks take
contiguous blocks of pairs (each block needs one halo frame: its first `prev`) and the only
exchange is ONE ncclAllGather of the packed Bloom bit arrays.  torch.distributed is used for the
rendezvous only (broadcast of the NCCL unique id, barriers, max-over-ranks timing).
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _cabi


def shard_pairs(pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of pair indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(pairs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_unique_id(dist, make_id) -> np.ndarray:
    """Rank 0 calls make_id() -> uint8[128]; every rank returns the same 128 bytes (torch.distributed broadcast)."""
    import torch
    ident = np.zeros(128, dtype=np.uint8)
    if dist.get_rank() == 0:
        ident[:] = make_id()
    t = torch.from_numpy(ident.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    return t.cpu().numpy().copy()


def agree_slot_bytes(dist, local_max_bits: int, align: int = 16) -> int:
    """Fixed all-gather slot: max over ranks of ceil(l/8), rounded up (every rank must use the same size)."""
    import torch
    t = torch.tensor([(int(local_max_bits) + 7) // 8], dtype=torch.int64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return (int(t.item()) + align - 1) // align * align


def _nccl_unique_id() -> np.ndarray:
    ident = np.zeros(128, dtype=np.uint8)
    _cabi.check(_cabi.lib().rbf_nccl_unique_id(_cabi.ptr(ident)))
    return ident


def init_nccl_from_torch(dist) -> None:
    """Create the library's NCCL communicator; the 128-byte unique id travels over torch.distributed."""
    ident = broadcast_unique_id(dist, _nccl_unique_id)
    _cabi.check(_cabi.lib().rbf_nccl_init(_cabi.ctx(), _cabi.ptr(ident), dist.get_rank(), dist.get_world_size()), _cabi.ctx())


class DeviceBuffer:
    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        self.nbytes = int(nbytes)
        _cabi.check(_cabi.lib().rbf_malloc(_cabi.ctx(), self.nbytes, C.byref(self.ptr)), _cabi.ctx())

    def to_host(self) -> np.ndarray:
        out = np.empty(self.nbytes, dtype=np.uint8)
        _cabi.check(_cabi.lib().rbf_memcpy_d2h(_cabi.ctx(), _cabi.ptr(out), self.ptr, self.nbytes), _cabi.ctx())
        return out

    def free(self):
        if self.ptr:
            _cabi.lib().rbf_free(_cabi.ctx(), self.ptr)
            self.ptr = None


def allgather_bitmaps(stream, pairs: int, slot_bytes: int, world: int):
    """One ncclAllGather of `pairs` fixed-size bitmap slots per rank -> DeviceBuffer [world][pairs][slot]."""
    send = DeviceBuffer(slot_bytes * pairs)
    recv = DeviceBuffer(slot_bytes * pairs * world)
    _cabi.check(_cabi.lib().rbf_stream_allgather_bitmaps(stream._h, pairs, slot_bytes, send.ptr, recv.ptr), _cabi.ctx())
    return send, recv
