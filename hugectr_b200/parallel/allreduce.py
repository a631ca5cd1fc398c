"""Dense gradient exchange: in-place all-reduce of the single flat wgrad buffer.

Reference: ExchangeWgrad / AllReduceInPlaceComm (HugeCTR/src/exchange_wgrad.cpp:22-92,
HugeCTR/src/collectives/all_reduce_comm.cu:86-426).  Back-ends:
  NCCL     torch.distributed all_reduce (baseline; NVLS inside NCCL on NVSwitch)
  OneShot  custom P2P kernel over the symmetric heap: every rank reduces its 1/N slice by loading
           16-byte lines from all peers, then all-gathers by peer stores (csrc/p2p.cu)
Either back-end runs bucket by bucket on a communication stream while the backward pass is still
producing the earlier layers' gradients (begin_step / layer_done / finish_step).
"""
from __future__ import annotations

import torch

from ..enums import AllReduceAlgo


def advance_final(final_lo: int, pending):
    """Move ``final_lo`` down over finished ranges that touch it; returns (final_lo, still pending)."""
    moved = True
    while moved:
        moved = False
        for (a, b) in pending:
            if a < final_lo <= b:
                final_lo = a
                moved = True
        pending = [(a, b) for (a, b) in pending if a < final_lo]
    return final_lo, pending


class ExchangeWgrad:
    def __init__(self, comm, wgrad: torch.Tensor, algo: AllReduceAlgo = AllReduceAlgo.NCCL):
        self.comm = comm
        self.wgrad = wgrad
        self.algo = algo
        self.extra = []     # grouped all-reduce: extra buffers (DP embedding wgrads)
        self._p2p = None
        if comm.world_size > 1 and algo in (AllReduceAlgo.OneShot, AllReduceAlgo.TwoShot) \
                and comm.p2p_available and not getattr(comm, "emulated", False):
            # (emulated ranks share one device: a spinning kernel and a device-wide sync of another rank
            # thread could deadlock; the kernel has its own emulated-rank test)
            from .p2p import P2PAllReduce
            self._p2p = P2PAllReduce(comm, wgrad)

    def register_extra(self, t: torch.Tensor):
        self.extra.append(t)

    # ---- bucketed overlap with backward (F3): the flat wgrad is produced back to front (bprop
    # visits layers in reverse creation order), so [lo, hi) ranges become final incrementally
    def begin_step(self, after_bucket=None):
        """``after_bucket(lo, hi)``: optional callback run on the communication stream right after
        the bucket's all-reduce (the fused dense optimizer of that parameter range), so neither the
        reduction nor the update of the big upper layers sits at the tail of the step."""
        self._done_lo = self.wgrad.numel()      # everything in [_done_lo, end) has been flushed
        self._final_lo = self.wgrad.numel()     # everything in [_final_lo, end) is final (bprop done)
        self._pending = []                      # finished layer ranges below _final_lo (out of order)
        self._after = after_bucket
        if self.wgrad.is_cuda and not hasattr(self, "_stream"):
            import os
            default = "-3" if (os.environ.get("HCTR_STEP_SCHEDULE", "") == "aggressive" or
                               (os.environ.get("HCTR_STEP_SCHEDULE", "") != "safe" and self.comm.world_size <= 4)) else "0"
            prio = int(os.environ.get("HCTR_PRIO_COMM", default))
            self._stream = torch.cuda.Stream(priority=prio)       # all-reduces, back to back
            self._opt_stream = torch.cuda.Stream(priority=prio)   # per-bucket optimizer slices, behind their bucket

    def layer_done(self, lo: int, hi: int = None, bucket_elems: int = 2 << 20):
        """called after a trainable layer's bprop; [lo, hi) = arena range of its parameters.

        Layers need not finish back to front (the overlapped step runs the embedding-dependent "top"
        pass before the "bottom" pass, and a bottom layer may have been declared after a top layer):
        a range is flushed only when every parameter above it is final, i.e. ``_final_lo`` moves down
        over CONTIGUOUS finished ranges (a range ends where the next parameter of the arena starts); out-of-order ranges wait in ``_pending``."""
        if not self.wgrad.is_cuda:
            return
        if hi is None:
            hi = self._final_lo
        self._final_lo, self._pending = advance_final(self._final_lo, self._pending + [(lo, hi)])
        if self._done_lo - self._final_lo >= bucket_elems:
            self._flush(self._final_lo)


    def _flush(self, lo: int):
        hi = self._done_lo
        if hi <= lo:
            return
        main = torch.cuda.current_stream()
        self._stream.wait_stream(main)
        with torch.cuda.stream(self._stream):
            if self.comm.world_size > 1:
                if self._p2p is not None:
                    self._p2p.run(lo, hi)
                else:
                    self.comm.all_reduce(self.wgrad[lo:hi])
        if self._after is not None:
            # the update of this bucket must not delay the next bucket's all-reduce
            self._opt_stream.wait_stream(self._stream)
            with torch.cuda.stream(self._opt_stream):
                self._after(lo, hi)
        self._done_lo = lo

    def finish_step(self):
        """all remaining ranges + join: wgrad is fully reduced (and, with ``after_bucket``, applied)
        on the current stream afterwards"""
        if not self.wgrad.is_cuda or not hasattr(self, "_done_lo"):
            self.allreduce()
            if getattr(self, "_after", None) is not None:
                self._after(0, self.wgrad.numel())
            return
        self._flush(0)
        torch.cuda.current_stream().wait_stream(self._stream)
        torch.cuda.current_stream().wait_stream(self._opt_stream)
        for t in self.extra:
            self.comm.all_reduce(t)

    def allreduce(self):
        if self.comm.world_size == 1:
            return
        if self._p2p is not None:
            self._p2p.run()
        else:
            self.comm.all_reduce(self.wgrad)
        for t in self.extra:
            self.comm.all_reduce(t)
