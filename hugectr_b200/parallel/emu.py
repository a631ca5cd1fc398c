"""In-process multi-rank fabric: N "ranks" as N Python threads sharing one device.

Every rank owns a CUDA stream (thread-local current stream) and an ``EmuComm`` with the ``Comm`` API.
Collectives rendezvous on the host (``threading.Barrier``); the symmetric heap is a list of local
allocations per rank, so ``peer_ptrs`` are ordinary device pointers and the SAME peer-memory kernels
(owner-side embedding forward / backward, device barrier, two-shot all-reduce) run unmodified -- the
loads and stores that cross NVLink on the real machine stay inside one GPU here.

Why it exists: the GPU test box of the driver has ONE GPU.  The peer kernels, the requester-side shard
split, the fused backward and whole-model multi-rank equivalence are exercised there through this fabric
(``tests/test_emu_ranks.py``); the same entry points run under torchrun on real multi-GPU boxes
(``tests/test_dist.py``).  On CPU tensors it doubles as a fast multi-rank test bed without process spawns.

Not a performance tool: ranks time-share one GPU and every collective synchronises the host.
"""
from __future__ import annotations

import threading
import traceback
from typing import Callable, List

import torch


class EmuFabric:
    def __init__(self, world: int, device):
        self.world = int(world)
        self.device = torch.device(device)
        self.bar = threading.Barrier(self.world)
        self.slots: List[object] = [None] * self.world
        self.allocs: List[List[torch.Tensor]] = [[] for _ in range(self.world)]
        self.device_barriers = 0

    def exchange(self, rank: int, obj):
        """all-gather of Python objects (references, not copies)"""
        self.slots[rank] = obj
        self.bar.wait()
        out = list(self.slots)
        self.bar.wait()
        return out


class EmuComm:
    """``Comm`` look-alike of one emulated rank (see hugectr_b200/parallel/comm.py for the contracts)."""

    def __init__(self, fabric: EmuFabric, rank: int, p2p: bool = True):
        self.fabric, self.rank, self.world_size = fabric, rank, fabric.world
        self.device = fabric.device
        self.group = None
        # p2p="force": fused data flow on CPU tensors too (logic tests of the peer-memory paths)
        self._p2p = (p2p == "force" or (bool(p2p) and self.device.type == "cuda")) and self.world_size > 1
        self.num_nodes = 1
        self.emulated = True

    # ---- helpers
    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def _xchg(self, obj):
        self._sync()
        return self.fabric.exchange(self.rank, obj)

    def _done(self):
        """nobody may overwrite a tensor another rank is still reading"""
        self._sync()
        self.fabric.bar.wait()

    # ---- collectives
    def all_reduce(self, t: torch.Tensor):
        if self.world_size > 1:
            parts = self._xchg(t)
            s = parts[0].clone()
            for p in parts[1:]:
                s += p
            self._done()
            t.copy_(s)
            self._done()
        return t

    def all_gather(self, out: torch.Tensor, inp: torch.Tensor):
        parts = self._xchg(inp.contiguous().view(-1))
        n = parts[0].numel()
        o = out.view(-1)
        for r, p in enumerate(parts):
            o[r * n:(r + 1) * n].copy_(p)
        self._done()

    def all_to_all(self, out: torch.Tensor, inp: torch.Tensor):
        parts = self._xchg(inp)
        for r, p in enumerate(parts):
            out[r].copy_(p[self.rank].to(out.dtype))
        self._done()

    def all_to_all_v(self, out, inp, out_splits, in_splits):
        offs = [0]
        for n in in_splits:
            offs.append(offs[-1] + int(n))
        parts = self._xchg((inp, offs))
        o = 0
        for r, (p, po) in enumerate(parts):
            n = int(out_splits[r])
            out[o:o + n].copy_(p[po[self.rank]:po[self.rank] + n].to(out.dtype))
            o += n
        self._done()

    def reduce_scatter(self, out: torch.Tensor, inp: torch.Tensor):
        parts = self._xchg(inp.contiguous().view(self.world_size, -1))
        s = parts[0][self.rank].clone().float()
        for p in parts[1:]:
            s += p[self.rank].float()
        self._done()
        out.view(-1).copy_(s.to(out.dtype))

    def broadcast(self, t: torch.Tensor, src: int = 0):
        parts = self._xchg(t)
        if self.rank != src:
            t.copy_(parts[src])
        self._done()
        return t

    def barrier(self):
        self._sync()
        self.fabric.bar.wait()

    def all_gather_object(self, obj):
        return self.fabric.exchange(self.rank, obj)

    def warmup(self):
        pass

    def set_topology(self, gpus_per_node: int):
        """logical nodes (rank = node * gpus_per_node + local id, like ``Comm.set_topology``): one intra-node and
        one inter-node emulated communicator per rank, so the hierarchical exchange runs on emulated ranks too"""
        L = int(gpus_per_node)
        if getattr(self, "_topo", None) == L:
            return
        self._topo = L
        if self.world_size == 1 or L <= 0 or L >= self.world_size or self.world_size % L:
            self.num_nodes, self.local_size = 1, self.world_size
            return
        nodes = self.world_size // L
        node, lid = self.rank // L, self.rank % L
        # the first member of every group creates its fabric; everybody learns all of them
        made = (EmuFabric(L, self.device) if lid == 0 else None, EmuFabric(nodes, self.device) if node == 0 else None)
        allf = self.fabric.exchange(self.rank, made)
        self.intra = EmuComm(allf[node * L][0], lid, p2p=False)
        self.inter = EmuComm(allf[lid][1], node, p2p=False)
        self.num_nodes, self.local_size, self.node, self.local_id = nodes, L, node, lid

    def hier_all_to_all_sum(self, send: torch.Tensor) -> torch.Tensor:
        from .comm import Comm
        return Comm.hier_all_to_all_sum(self, send)

    def hier_all_gather(self, out: torch.Tensor, inp: torch.Tensor):
        from .comm import Comm
        return Comm.hier_all_gather(self, out, inp)

    # ---- symmetric heap look-alike (``comm.heap`` is the communicator itself)
    @property
    def p2p_available(self) -> bool:
        return self._p2p

    def disable_p2p(self):
        self._p2p = False

    @property
    def heap(self):
        return self

    def alloc(self, numel: int, dtype: torch.dtype) -> torch.Tensor:
        t = torch.zeros(max(int(numel), 1), dtype=dtype, device=self.device)
        self.fabric.allocs[self.rank].append(t)
        self.barrier()                      # collective, like SymmetricHeap.alloc
        return t[:numel]

    symm_alloc = alloc

    def peer_ptrs(self, t: torch.Tensor) -> List[int]:
        base = t.data_ptr()
        for i, a in enumerate(self.fabric.allocs[self.rank]):
            lo = a.data_ptr()
            if lo <= base < lo + a.numel() * a.element_size():
                off = base - lo
                if not t.is_cuda:
                    # CPU fabric: the "peer pointers" are the peers' tensors themselves (the reference
                    # implementations of the kernels take tensors), so the FUSED data flow -- routes,
                    # inbox layouts, barrier placement -- can be exercised without a GPU
                    e0 = off // a.element_size()
                    return [self.fabric.allocs[r][i].view(-1)[e0:e0 + t.numel()].view(t.shape)
                            for r in range(self.world_size)]
                return [self.fabric.allocs[r][i].data_ptr() + off for r in range(self.world_size)]
        raise KeyError("tensor is not part of the (emulated) symmetric heap")

    def barrier_device(self):
        """the device barrier of the real heap orders kernels across GPUs; here: drain my stream,
        then a host rendezvous (every rank's earlier kernels are complete before anyone continues)"""
        if self.world_size > 1:
            self.fabric.device_barriers += 1
            self._sync()
            self.fabric.bar.wait()

    def shutdown(self, destroy_process_group: bool = False):
        pass


def run_ranks(world: int, fn: Callable[[EmuComm], object], device=None, p2p: bool = True, timeout: float = 600.0):
    """Run ``fn(comm)`` on ``world`` emulated ranks (threads); returns the list of results, re-raises
    the first failure (the other ranks are released through a broken barrier)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
            else torch.device("cpu")
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    fab = EmuFabric(world, device)
    res: List[object] = [None] * world
    errs: List[object] = [None] * world

    def body(r):
        try:
            comm = EmuComm(fab, r, p2p)
            if fab.device.type == "cuda":
                torch.cuda.set_device(fab.device)
                with torch.cuda.stream(torch.cuda.Stream(fab.device)):
                    res[r] = fn(comm)
                    torch.cuda.current_stream().synchronize()
            else:
                res[r] = fn(comm)
        except BaseException as e:      # noqa: BLE001 -- reported to the caller below
            errs[r] = (e, traceback.format_exc())
            fab.bar.abort()

    ths = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout)
        if t.is_alive():
            fab.bar.abort()
            raise TimeoutError("emulated rank did not finish")
    real = [e for e in errs if e is not None and not isinstance(e[0], threading.BrokenBarrierError)]
    if real:
        raise RuntimeError("emulated rank failed:\n" + real[0][1]) from real[0][0]
    if any(e is not None for e in errs):
        raise RuntimeError("emulated ranks aborted:\n" + next(e for e in errs if e is not None)[1])
    return res
