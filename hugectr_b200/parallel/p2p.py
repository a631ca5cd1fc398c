"""P2P collectives built on the symmetric heap (csrc/p2p.cu)."""
from __future__ import annotations

import ctypes as C

import torch

from ..ops import dense as D
from . import symm as S


class P2PAllReduce:
    """In-place two-shot all-reduce of a flat fp32 buffer that lives in the symmetric heap."""

    def __init__(self, comm, buf: torch.Tensor, blocks: int = 64):
        self.comm = comm
        heap = comm.heap
        self.buf = buf
        self._peer_list = heap.peer_ptrs(buf)
        self.peers = S.ptr_array(self._peer_list)
        self.flags = heap.alloc(64, torch.int32)
        self.flag_ptrs = S.ptr_array(heap.peer_ptrs(self.flags))
        self.epoch = torch.zeros(1, dtype=torch.int32, device=comm.device)
        self.gate = torch.zeros(2, dtype=torch.int32, device=comm.device)
        self.gate_epoch = torch.zeros(1, dtype=torch.int32, device=comm.device)
        self.blocks = blocks
        assert buf.numel() % (4 * comm.world_size) == 0, "pad the buffer to 4*world elements"

    def run(self, lo: int = 0, hi: int = None):
        """all-reduce elements [lo, hi) (both multiples of 4*world) on the current stream"""
        hi = self.buf.numel() if hi is None else hi
        if hi <= lo:
            return
        peers = self.peers
        if lo:
            peers = S.ptr_array([int(q) + lo * 4 for q in self._peer_list])
        # small ranges: fewer blocks (the kernel holds an in-kernel grid barrier, so every block must
        # become resident next to whatever else is running before any data moves)
        per_rank_vec4 = (hi - lo) // (4 * self.comm.world_size)
        blocks = max(4, min(self.blocks, (per_rank_vec4 + 1023) // 1024))
        rc = S.lib().hctr_allreduce_twoshot(
            peers, self.flag_ptrs, self.epoch.data_ptr(), self.gate.data_ptr(),
            self.gate_epoch.data_ptr(), hi - lo, self.comm.rank, self.comm.world_size,
            blocks, torch.cuda.current_stream(self.comm.device).cuda_stream)
        if rc:
            raise RuntimeError(f"allreduce_twoshot failed rc={rc}")
        D._count()


def peer_pull(comm, src_ptrs, dst_tensors, nbytes_each, blocks: int = 64):
    """dst[r] (local) <- src_ptrs[r] (peer), 16-byte vectorised."""
    n = len(src_ptrs)
    src = S.ptr_array(src_ptrs)
    dst = S.ptr_array([t.data_ptr() for t in dst_tensors])
    n16 = (C.c_longlong * n)(*[b // 16 for b in nbytes_each])
    rc = S.lib().hctr_peer_pull(src, dst, n16, n, blocks,
                                torch.cuda.current_stream(comm.device).cuda_stream)
    if rc:
        raise RuntimeError("peer_pull failed")
    D._count()
