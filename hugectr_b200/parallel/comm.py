"""Process-group plumbing: one process per GPU, torch.distributed (NCCL / gloo) for bootstrap and
baseline collectives, a peer-mapped symmetric heap + device-side barrier for the fused kernels.

Replaces ResourceManagerCore / GPUResource / CollectiveManager of the reference
(HugeCTR/src/resource_managers/resource_manager_core.cpp:36-259, HugeCTR/src/gpu_resource.cpp):
streams/handles are PyTorch's, NCCL communicators are torch.distributed's, the all2all warm-up is
``warmup()``, P2P enablement is the symmetric heap's IPC mapping.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


class DeviceMap:
    """vvgpu -> local/global ids (HugeCTR/include/device_map.hpp:76-115)."""

    def __init__(self, vvgpu, layout="LocalFirst", my_node: int = 0):
        self.vvgpu = [list(v) for v in vvgpu]
        self.layout = getattr(layout, "name", layout)
        self.my_node = my_node
        self.num_nodes = len(self.vvgpu)
        n_local = [len(v) for v in self.vvgpu]
        self.global_ids = {}
        if self.layout == "NodeFirst":
            assert len(set(n_local)) == 1, "NodeFirst needs the same GPU count on every node"
            for node, devs in enumerate(self.vvgpu):
                for li, d in enumerate(devs):
                    self.global_ids[(node, li)] = li * self.num_nodes + node
        else:
            g = 0
            for node, devs in enumerate(self.vvgpu):
                for li, d in enumerate(devs):
                    self.global_ids[(node, li)] = g
                    g += 1
        self.total = sum(n_local)

    def get_global_id(self, local_id: int, node: Optional[int] = None) -> int:
        return self.global_ids[(self.my_node if node is None else node, local_id)]

    def get_local_devices(self, node: Optional[int] = None) -> List[int]:
        return list(self.vvgpu[self.my_node if node is None else node])

    def get_pid(self, global_id: int) -> int:
        for (node, li), g in self.global_ids.items():
            if g == global_id:
                return node
        raise KeyError(global_id)

    def size(self) -> int:
        return self.total


class Comm:
    """Communicator facade. world_size == 1 needs no process group."""

    def __init__(self, device: torch.device, group=None):
        self.device = device
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world_size = dist.get_world_size(group)
        else:
            self.rank, self.world_size = 0, 1
        self._heap = None
        self._p2p = None

    @staticmethod
    def single(device) -> "Comm":
        """A world_size == 1 communicator even inside a multi-rank job (tests, oracles)."""
        c = Comm.__new__(Comm)
        c.device, c.group, c.rank, c.world_size, c._heap, c._p2p = device, None, 0, 1, None, False
        return c

    # ---- bootstrap
    @staticmethod
    def init_from_env(device_type: Optional[str] = None) -> "Comm":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        use_cuda = torch.cuda.is_available() if device_type is None else device_type == "cuda"
        if use_cuda:
            lr = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(lr)
            device = torch.device("cuda", lr)
        else:
            device = torch.device("cpu")
        if world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            kw = {}
            if use_cuda:
                kw["device_id"] = device
            t = float(os.environ.get("HCTR_STEP_TIMEOUT", "0") or 0)
            if t > 0:       # collectives give up with the step watchdog (utils/watchdog.py)
                import datetime
                kw["timeout"] = datetime.timedelta(seconds=max(t, 10.0))
            dist.init_process_group("nccl" if use_cuda else "gloo", **kw)
        return Comm(device)

    # ---- baseline collectives (NCCL / gloo)
    def all_reduce(self, t: torch.Tensor):
        if self.world_size > 1:
            dist.all_reduce(t, group=self.group)
        return t

    def all_gather(self, out: torch.Tensor, inp: torch.Tensor):
        if self.world_size == 1:
            out.view(-1)[:inp.numel()].copy_(inp.view(-1))
        else:
            dist.all_gather_into_tensor(out.view(-1), inp.contiguous().view(-1), group=self.group)

    def all_to_all(self, out: torch.Tensor, inp: torch.Tensor):
        if self.world_size == 1:
            out.copy_(inp)
        elif inp.is_cuda:
            dist.all_to_all_single(out.view(-1), inp.contiguous().view(-1), group=self.group)
        else:  # gloo has no all_to_all_single for every dtype: emulate with all_gather
            n = inp.shape[0]
            gathered = [torch.empty_like(inp) for _ in range(n)]
            src = inp.float() if inp.dtype == torch.bfloat16 else inp
            gl = [torch.empty_like(src) for _ in range(n)]
            dist.all_gather(gl, src.contiguous(), group=self.group)
            for j in range(n):
                out[j].copy_(gl[j][self.rank].to(out.dtype))

    def all_to_all_v(self, out: torch.Tensor, inp: torch.Tensor, out_splits, in_splits):
        """variable-size all-to-all of flat tensors: ``inp`` = concatenation of the chunks for ranks
        0..n-1 (``in_splits`` elements each), ``out`` = concatenation of the chunks received from them
        (``out_splits``).  NCCL: one ``all_to_all_single`` with split sizes; gloo: padded all-gather."""
        if self.world_size == 1:
            out.copy_(inp)
        elif inp.is_cuda:
            dist.all_to_all_single(out, inp.contiguous(), [int(x) for x in out_splits],
                                   [int(x) for x in in_splits], group=self.group)
        else:
            # every rank broadcasts its chunks padded to its own widest one; receivers keep their row
            src = inp.float() if inp.dtype == torch.bfloat16 else inp
            mx = torch.tensor([int(max(max(in_splits), 1))])
            widths = [torch.zeros_like(mx) for _ in range(self.world_size)]
            dist.all_gather(widths, mx, group=self.group)
            o_out = 0
            for r in range(self.world_size):
                blk = torch.zeros(self.world_size, int(widths[r]), dtype=src.dtype)
                if r == self.rank:
                    o = 0
                    for dst, n in enumerate(in_splits):
                        blk[dst, :n] = src[o:o + n]
                        o += n
                dist.broadcast(blk, r, group=self.group)
                n = int(out_splits[r])
                out[o_out:o_out + n].copy_(blk[self.rank, :n].to(out.dtype))
                o_out += n

    def reduce_scatter(self, out: torch.Tensor, inp: torch.Tensor):
        if self.world_size == 1:
            out.copy_(inp.view_as(out))
        elif inp.is_cuda:
            dist.reduce_scatter_tensor(out.view(-1), inp.contiguous().view(-1), group=self.group)
        else:
            tmp = inp.clone().float()
            dist.all_reduce(tmp, group=self.group)
            n = out.numel()
            out.view(-1).copy_(tmp.view(-1)[self.rank * n:(self.rank + 1) * n].to(out.dtype))

    def broadcast(self, t: torch.Tensor, src: int = 0):
        if self.world_size > 1:
            dist.broadcast(t, src, group=self.group)
        return t

    def barrier(self):
        if self.world_size > 1:
            dist.barrier(group=self.group)

    def all_gather_object(self, obj):
        if self.world_size == 1:
            return [obj]
        out = [None] * self.world_size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def warmup(self):
        """resource_manager_core.cpp:36-75 all2all warm-up."""
        if self.world_size > 1:
            t = torch.zeros(self.world_size, dtype=torch.int64, device=self.device)
            o = torch.zeros_like(t)
            self.all_to_all(o.view(self.world_size, 1), t.view(self.world_size, 1))
            self.all_reduce(t)

    # ---- node topology (hierarchical collectives)
    def set_topology(self, gpus_per_node: int):
        """Declare the node structure (global rank = node * gpus_per_node + local id, the LocalFirst
        layout of torchrun).  Builds one intra-node group per node and one inter-node group per local
        id; collective over ALL ranks (every rank creates every group in the same order)."""
        if getattr(self, "_topo", None) == gpus_per_node:
            return
        L = int(gpus_per_node)
        if self.world_size == 1 or L <= 0 or L >= self.world_size or self.world_size % L:
            self._topo, self.num_nodes, self.local_size = gpus_per_node, 1, self.world_size
            self.node, self.local_id, self.intra, self.inter = 0, self.rank, self, None
            return
        nodes = self.world_size // L
        self.num_nodes, self.local_size = nodes, L
        self.node, self.local_id = self.rank // L, self.rank % L
        intra = inter = None
        for n in range(nodes):
            g = dist.new_group([n * L + l for l in range(L)])
            if n == self.node:
                intra = g
        for l in range(L):
            g = dist.new_group([n * L + l for n in range(nodes)])
            if l == self.local_id:
                inter = g
        self.intra, self.inter = Comm(self.device, intra), Comm(self.device, inter)
        self._topo = gpus_per_node

    def hier_all_to_all_sum(self, send: torch.Tensor) -> torch.Tensor:
        """send[d] = my partial contribution to rank d ([world, n]); returns sum over all ranks of
        their contribution to me ([n]).  Two stages like the reference's hierarchical model-parallel
        embedding (hier_model_parallel_embedding.cpp:183-230): partial sums that leave the node for
        the same destination are reduced inside the node first (reduce-scatter onto the local rank
        with the destination's local id), then ONE exchange per node pair between same-local-id
        ranks."""
        nodes, L = self.num_nodes, self.local_size
        n = send.shape[-1]
        by_local = send.view(nodes, L, n).transpose(0, 1).contiguous()      # [L, nodes, n]
        red = torch.empty(nodes, n, dtype=send.dtype, device=send.device)
        self.intra.reduce_scatter(red, by_local)                            # sum over my node
        recv = torch.empty_like(red)
        self.inter.all_to_all(recv, red)                                    # chunk m <-> node m
        return recv.float().sum(0).to(send.dtype)

    def hier_all_gather(self, out: torch.Tensor, inp: torch.Tensor):
        """out [world, n] <- every rank's inp [n], gathered across nodes first (same local id), then
        inside the node"""
        nodes, L = self.num_nodes, self.local_size
        n = inp.numel()
        stage1 = torch.empty(nodes, n, dtype=inp.dtype, device=inp.device)
        self.inter.all_gather(stage1, inp)
        stage2 = torch.empty(L, nodes, n, dtype=inp.dtype, device=inp.device)
        self.intra.all_gather(stage2, stage1)
        out.view(nodes, L, n).copy_(stage2.transpose(0, 1))

    # ---- symmetric heap / P2P
    @property
    def p2p_available(self) -> bool:
        if self._p2p is None:
            ok = False
            if self.world_size > 1 and self.device.type == "cuda" and \
                    os.environ.get("HCTR_DISABLE_P2P", "0") == "0":
                try:
                    from .symm import SymmetricHeap
                    self._heap = SymmetricHeap(self)
                    ok = True
                except Exception as e:  # pragma: no cover
                    from ..utils import logger
                    logger.warning(f"P2P symmetric heap unavailable ({e}); using NCCL collectives")
            self._p2p = ok
        return self._p2p

    @property
    def heap(self):
        if not self.p2p_available:
            raise RuntimeError("symmetric heap not available")
        return self._heap

    def symm_alloc(self, numel: int, dtype: torch.dtype) -> torch.Tensor:
        return self.heap.alloc(numel, dtype)

    def peer_ptrs(self, t: torch.Tensor) -> List[int]:
        return self.heap.peer_ptrs(t)

    def disable_p2p(self):
        """Force the NCCL / gloo collective paths on this communicator (the stand-in baseline arm)."""
        if self._heap is not None:
            self._heap.close()
            self._heap = None
        self._p2p = False

    def shutdown(self, destroy_process_group: bool = True):
        """Orderly teardown (collective): close the symmetric heap, then the process group.  After
        this the interpreter exits through its normal finalizers."""
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        if self._heap is not None:
            self._heap.close()
            self._heap = None
            self._p2p = False
        if destroy_process_group and self.group is None and dist.is_available() and dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()

    def barrier_device(self):
        """Device-side all-GPU barrier on the current stream (no host sync, graph capturable)."""
        if self.world_size > 1:
            self.heap.barrier()
