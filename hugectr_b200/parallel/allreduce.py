"""Dense gradient exchange: in-place all-reduce of the single flat wgrad buffer.

Reference: ExchangeWgrad / AllReduceInPlaceComm (HugeCTR/src/exchange_wgrad.cpp:22-92,
HugeCTR/src/collectives/all_reduce_comm.cu:86-426).  Back-ends:
  NCCL     torch.distributed all_reduce (baseline; NVLS inside NCCL on NVSwitch)
  OneShot  custom P2P kernel over the symmetric heap: every rank reduces its 1/N slice by loading
           16-byte lines from all peers, then all-gathers by peer stores (csrc/p2p.cu)
"""
from __future__ import annotations

import torch

from ..enums import AllReduceAlgo


class ExchangeWgrad:
    def __init__(self, comm, wgrad: torch.Tensor, algo: AllReduceAlgo = AllReduceAlgo.NCCL):
        self.comm = comm
        self.wgrad = wgrad
        self.algo = algo
        self.extra = []     # grouped all-reduce: extra buffers (DP embedding wgrads)
        self._p2p = None
        if comm.world_size > 1 and algo in (AllReduceAlgo.OneShot, AllReduceAlgo.TwoShot) \
                and comm.p2p_available:
            from .p2p import P2PAllReduce
            self._p2p = P2PAllReduce(comm, wgrad)

    def register_extra(self, t: torch.Tensor):
        self.extra.append(t)

    def allreduce(self):
        if self.comm.world_size == 1:
            return
        if self._p2p is not None:
            self._p2p.run()
        else:
            self.comm.all_reduce(self.wgrad)
        for t in self.extra:
            self.comm.all_reduce(t)
