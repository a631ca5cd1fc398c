"""Symmetric peer-mapped heap: every rank allocates the same buffers with cudaMalloc, exchanges CUDA
IPC handles through torch.distributed and maps every peer's copy, so kernels can ld/st peer HBM
directly over NVLink.  The generalisation of the reference's cudaDeviceEnablePeerAccess pointer
tables (HugeCTR/src/resource_managers/resource_manager_core.cpp:77-106) to one process per GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch

from .. import _native


class _RawCuda:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1",
                                         "data": (ptr, False), "version": 3}


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = _native.cuda_lib()
        vp, i, ll = C.c_void_p, C.c_int, C.c_longlong
        l.hctr_ipc_alloc.argtypes = [ll]
        l.hctr_ipc_alloc.restype = vp
        l.hctr_ipc_free.argtypes = [vp]
        l.hctr_ipc_get_handle.argtypes = [vp, vp]
        l.hctr_ipc_open.argtypes = [vp]
        l.hctr_ipc_open.restype = vp
        l.hctr_ipc_close.argtypes = [vp]
        l.hctr_peer_barrier.argtypes = [C.POINTER(vp), vp, i, i, vp]
        l.hctr_allreduce_twoshot.argtypes = [C.POINTER(vp), C.POINTER(vp), vp, vp, vp, ll, i, i, i, vp]
        l.hctr_peer_pull.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(ll), i, i, vp]
        l.hctr_last_cuda_error.restype = C.c_char_p
        _lib = l
    return _lib


def ptr_array(ptrs: List[int]):
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


class SymmetricHeap:
    def __init__(self, comm):
        self.comm = comm
        self.rank, self.world = comm.rank, comm.world_size
        self.device = comm.device
        torch.cuda.set_device(self.device)
        self._allocs: Dict[int, dict] = {}
        # barrier state: flags [16] uint32 per rank (peer mapped) + private epoch counter
        self.flags = self.alloc(64, torch.int32)
        self.flag_ptrs = ptr_array(self.peer_ptrs(self.flags))
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.comm.barrier()

    def alloc(self, numel: int, dtype: torch.dtype) -> torch.Tensor:
        """Collective: every rank must call with the same arguments in the same order."""
        esz = torch.empty(0, dtype=dtype).element_size()
        nbytes = (numel * esz + 255) // 256 * 256
        # Every rank walks through the same collective sequence even when a local step fails, and the
        # outcome is agreed on before anyone raises: either all ranks get the buffer or all ranks see
        # the error (and fall back to NCCL collectives together) -- never a mix that would deadlock.
        err = None
        p = lib().hctr_ipc_alloc(nbytes)
        h = C.create_string_buffer(64)
        if not p:
            err = "cudaMalloc failed for the symmetric heap: %s" % lib().hctr_last_cuda_error().decode()
        elif lib().hctr_ipc_get_handle(p, h) != 0:
            err = "cudaIpcGetMemHandle failed"
        handles = self.comm.all_gather_object(None if err else bytes(h.raw))
        peers = []
        if err is None and all(hb is not None for hb in handles):
            for r, hb in enumerate(handles):
                if r == self.rank:
                    peers.append(p)
                    continue
                q = lib().hctr_ipc_open(C.create_string_buffer(hb, 64))
                if not q:
                    err = "cudaIpcOpenMemHandle failed: %s" % lib().hctr_last_cuda_error().decode()
                    break
                peers.append(q)
        errs = self.comm.all_gather_object(err)
        if any(e is not None for e in errs) or any(hb is None for hb in handles):
            for r, q in enumerate(peers):
                if r != self.rank:
                    lib().hctr_ipc_close(q)
            if p:
                lib().hctr_ipc_free(p)
            bad = [(r, e) for r, e in enumerate(errs) if e is not None]
            raise RuntimeError("symmetric heap allocation failed on rank(s) %s" %
                               ", ".join(f"{r}: {e}" for r, e in bad))
        t = torch.as_tensor(_RawCuda(p, nbytes), device=self.device).view(dtype)[:numel]
        self._allocs[p] = {"peers": peers, "nbytes": nbytes, "tensor": t}
        return t

    def peer_ptrs(self, t: torch.Tensor) -> List[int]:
        base = t.data_ptr()
        for p, info in self._allocs.items():
            if p <= base < p + info["nbytes"]:
                off = base - p
                return [q + off for q in info["peers"]]
        raise KeyError("tensor is not part of the symmetric heap")

    def close(self):
        """Collective, idempotent: unmap every peer's buffers and free the local ones.  Called from
        ``Comm.shutdown`` after the last kernel that touches peer memory has completed on EVERY rank
        (device sync + process-group barrier on both sides), so no rank unmaps memory a peer still
        reads and interpreter exit never waits on a mapping whose exporter is already gone."""
        if getattr(self, "_closed", False):
            return
        self._closed = True
        torch.cuda.synchronize(self.device)
        self.comm.barrier()
        for p, info in self._allocs.items():
            for r, q in enumerate(info["peers"]):
                if r != self.rank and q:
                    lib().hctr_ipc_close(q)
            info["peers"] = []
        torch.cuda.synchronize(self.device)
        self.comm.barrier()                  # every importer has closed before any exporter frees
        for p, info in self._allocs.items():
            info["tensor"] = None
            lib().hctr_ipc_free(p)
        self._allocs = {}

    def barrier(self):
        rc = lib().hctr_peer_barrier(self.flag_ptrs, self.epoch.data_ptr(), self.rank, self.world,
                                     torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            raise RuntimeError("peer barrier launch failed")
        from ..ops import dense as D
        D._count()
