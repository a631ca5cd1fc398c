"""Embedding kernels: ctypes bindings for csrc/embedding.cu + PyTorch reference implementations.

The reference implementations define the semantics (they are the CPU path and the test oracle):
  forward   pooled[src, lookup, sample] = combiner( table[row_off + key // k] for owned keys )
  backward  per unique arena row: g = sum of bucket gradients (scaled 1/nnz for mean);
            then the sparse optimizer of Appendix A.3 on that row.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List

import torch

from .. import _native
from ..enums import Optimizer_t
from ..ops import dense as D

MAX_RANKS = 16
OPT_CODE = {Optimizer_t.SGD: 0, Optimizer_t.AdaGrad: 1, Optimizer_t.Adam: 2, Optimizer_t.Ftrl: 3,
            Optimizer_t.MomentumSGD: 4, Optimizer_t.Nesterov: 5, Optimizer_t.RMSProp: 6}


class CEmbLookup(C.Structure):
    _fields_ = [("table_row_off", C.c_longlong), ("key_off", C.c_longlong),
                ("nnz_off", C.c_longlong), ("out_off", C.c_longlong), ("grad_off", C.c_longlong),
                ("hotness", C.c_int), ("key_stride", C.c_int), ("num_shards", C.c_int),
                ("shard_idx", C.c_int), ("out_stride", C.c_int), ("grad_stride", C.c_int),
                ("combiner", C.c_int), ("ev_size", C.c_int), ("rows", C.c_int), ("pad_", C.c_int),
                ("pair_off", C.c_longlong)]


class CEmbParams(C.Structure):
    _fields_ = [("num_ranks", C.c_int), ("my_rank", C.c_int), ("batch", C.c_int),
                ("num_lookups", C.c_int), ("keys", C.c_void_p * MAX_RANKS),
                ("nnz", C.c_void_p * MAX_RANKS), ("out", C.c_void_p * MAX_RANKS),
                ("grad", C.c_void_p * MAX_RANKS), ("lookups", C.c_void_p), ("table", C.c_void_p),
                ("ev_size", C.c_int)]


class CUniqueTable(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("vals", C.c_void_p), ("counter", C.c_void_p),
                ("rows", C.c_void_p), ("slots", C.c_void_p), ("mask", C.c_uint),
                ("max_unique", C.c_uint)]


class CBwdIndex(C.Structure):
    _fields_ = [("count", C.c_void_p), ("offsets", C.c_void_p), ("block_sums", C.c_void_p),
                ("pair_uid", C.c_void_p), ("bucket_list", C.c_void_p), ("bucket_scale", C.c_void_p),
                ("heavy_items", C.c_void_p),
                ("heavy_count", C.c_void_p), ("heavy_scratch", C.c_void_p),
                ("heavy_ticket", C.c_void_p), ("heavy_slot", C.c_void_p),
                ("max_heavy_rows", C.c_uint), ("max_heavy_items", C.c_uint),
                ("light_list", C.c_void_p), ("medium_list", C.c_void_p), ("class_count", C.c_void_p)]


class COptHyper(C.Structure):
    _fields_ = [("lr_ptr", C.c_void_p), ("lr_scale", C.c_float), ("scaler", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
                ("lambda1", C.c_float), ("lambda2", C.c_float), ("ftrl_beta", C.c_float),
                ("momentum", C.c_float), ("initial_accu", C.c_float), ("step_ptr", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = _native.cuda_lib()
        vp, i, f, ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
        l.hctr_emb_forward.argtypes = [C.POINTER(CEmbParams), i, i, i, vp]
        l.hctr_emb_backward_accum.argtypes = [C.POINTER(CEmbParams), C.POINTER(CUniqueTable), vp, f,
                                              i, i, i, i, vp]
        l.hctr_emb_update.argtypes = [vp, vp, vp, vp, C.POINTER(CUniqueTable), i, i, i,
                                      C.POINTER(COptHyper), vp, i, vp]
        l.hctr_emb_gather_rows.argtypes = [vp, vp, vp, ll, i, i, vp]
        l.hctr_emb_bwd_index.argtypes = [C.POINTER(CEmbParams), C.POINTER(CUniqueTable),
                                         C.POINTER(CBwdIndex), ll, i, vp]
        l.hctr_emb_bwd_reduce_update.argtypes = [C.POINTER(CEmbParams), C.POINTER(CUniqueTable),
                                                 C.POINTER(CBwdIndex), vp, vp, i, i,
                                                 C.POINTER(COptHyper), f, i, vp, i, vp]
        for n in ("hctr_emb_forward", "hctr_emb_backward_accum", "hctr_emb_update",
                  "hctr_emb_gather_rows", "hctr_emb_bwd_index", "hctr_emb_bwd_reduce_update"):
            getattr(l, n).restype = i
        # ABI self-check: ctypes mirrors vs the structs the kernels were compiled with
        if hasattr(l, "hctr_abi_sizes_emb"):
            sz = (C.c_int * 8)()
            l.hctr_abi_sizes_emb(sz)
            mine = [C.sizeof(CEmbLookup), C.sizeof(CEmbParams), C.sizeof(CUniqueTable),
                    C.sizeof(CBwdIndex), C.sizeof(COptHyper)]
            if list(sz[:5]) != mine:
                raise RuntimeError(f"libhctr_cuda.so struct layout {list(sz[:5])} != python mirrors {mine}: "
                                   "rebuild the library (python -m hugectr_b200._native)")
        _lib = l
    return _lib


@dataclass
class LookupDesc:
    """Python mirror of EmbLookup (one owner-side lookup)."""
    table_row_off: int
    key_off: int
    out_off: int
    grad_off: int
    hotness: int
    key_stride: int
    num_shards: int
    shard_idx: int
    out_stride: int
    grad_stride: int
    combiner: int
    ev_size: int
    rows: int
    nnz_off: int = -1
    pair_off: int = 0

    def to_c(self) -> CEmbLookup:
        return CEmbLookup(self.table_row_off, self.key_off, self.nnz_off, self.out_off,
                          self.grad_off, self.hotness, self.key_stride, self.num_shards,
                          self.shard_idx, self.out_stride, self.grad_stride, self.combiner,
                          self.ev_size, self.rows, 0, self.pair_off)


def lookups_to_device(lookups: List[LookupDesc], device) -> torch.Tensor:
    arr = (CEmbLookup * max(1, len(lookups)))()
    for i, l in enumerate(lookups):
        arr[i] = l.to_c()
    raw = bytes(arr)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    return t.to(device)


def _st(dev):
    return torch.cuda.current_stream(dev).cuda_stream


# ----------------------------------------------------------------------------- shard split
def shard_split(key_slab, key_off, batch, hotness, k, split_off, nnz_slab, nnz_off):
    """Rewrite the bag ``key_slab[key_off : key_off + batch*hotness]`` ([batch, hotness], -1 = empty) of
    a table that is row-sharded k ways into k compacted lists of local row indices (key // k for the
    keys with key % k == j, original order, -1 padded) at ``split_off + j*batch*hotness`` and their
    lengths at ``nnz_slab[nnz_off + j*batch : ...]``."""
    if key_slab.is_cuda:
        l = lib()
        if not hasattr(l, "_split_ready"):
            l.hctr_emb_shard_split.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.c_void_p]
            l.hctr_emb_shard_split.restype = C.c_int
            l._split_ready = True
        esz = key_slab.element_size()
        rc = l.hctr_emb_shard_split(key_slab.data_ptr() + key_off * esz,
                                    key_slab.data_ptr() + split_off * esz,
                                    nnz_slab.data_ptr() + nnz_off * 4, batch, hotness, k, esz,
                                    _st(key_slab.device))
        if rc:
            raise RuntimeError("hctr_emb_shard_split failed")
        D._count()
        return
    keys = key_slab[key_off:key_off + batch * hotness].view(batch, hotness).long()
    out = key_slab[split_off:split_off + k * batch * hotness].view(k, batch, hotness)
    cnts = nnz_slab[nnz_off:nnz_off + k * batch].view(k, batch)
    pos = torch.arange(hotness).view(1, -1)
    for j in range(k):
        m = (keys >= 0) & (keys % k == j)
        order = torch.argsort((~m).to(torch.int8), dim=1, stable=True)
        rows = torch.div(keys, k, rounding_mode="floor").gather(1, order)
        cnt = m.sum(1)
        out[j].copy_(torch.where(pos < cnt.view(-1, 1), rows, torch.full_like(rows, -1)).to(out.dtype))
        cnts[j].copy_(cnt.to(torch.int32))


# ----------------------------------------------------------------------------- dispatch (peer stores)
class CDispatchRoute(C.Structure):
    _fields_ = [("src_off", C.c_longlong), ("dst_off", C.c_longlong), ("nnz_off", C.c_longlong),
                ("rows", C.c_int), ("row_elems", C.c_int), ("src_stride", C.c_int),
                ("dst_stride", C.c_int), ("dst_rank", C.c_int), ("kind", C.c_int), ("k", C.c_int),
                ("shard", C.c_int)]


@dataclass
class Route:
    """One block a rank sends to one destination rank's inbox (csrc/emb_dispatch.cu).  kind 0: 2-D
    copy of [rows, row_elems]; kind 1: key-bag split for shard ``shard`` of ``k`` (+ list lengths)."""
    src_off: int
    dst_off: int
    rows: int
    row_elems: int
    src_stride: int
    dst_stride: int
    dst_rank: int
    kind: int = 0
    k: int = 1
    shard: int = 0
    nnz_off: int = 0


def routes_to_device(routes: List[Route], device) -> torch.Tensor:
    arr = (CDispatchRoute * max(1, len(routes)))()
    for i, r in enumerate(routes):
        arr[i] = CDispatchRoute(r.src_off, r.dst_off, r.nnz_off, r.rows, r.row_elems, r.src_stride,
                                r.dst_stride, r.dst_rank, r.kind, r.k, r.shard)
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().to(device)


def dispatch(src: torch.Tensor, routes: List[Route], routes_dev, dst_bufs, nnz_bufs=None, blocks_x: int = 8):
    """Scatter blocks of the local 1-D buffer ``src`` into the per-rank inboxes ``dst_bufs`` (tensor or
    int peer pointer per destination rank; element type of ``src``).  Posted peer stores on CUDA."""
    if not routes:
        return
    if src.is_cuda:
        l = lib()
        if not hasattr(l, "_dispatch_ready"):
            l.hctr_emb_dispatch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]
            l.hctr_emb_dispatch.restype = C.c_int
            if l.hctr_abi_size_dispatch_route() != C.sizeof(CDispatchRoute):
                raise RuntimeError("libhctr_cuda.so DispatchRoute layout differs from the python mirror: rebuild")
            l._dispatch_ready = True
        n = len(dst_bufs)
        dp = (C.c_void_p * n)(*[d if isinstance(d, int) else d.data_ptr() for d in dst_bufs])
        np_ = None
        if nnz_bufs is not None:
            np_ = (C.c_void_p * n)(*[d if isinstance(d, int) else d.data_ptr() for d in nnz_bufs])
        rc = l.hctr_emb_dispatch(src.data_ptr(), routes_dev.data_ptr(), len(routes), dp, np_, n,
                                 src.element_size(), blocks_x, _st(src.device))
        if rc:
            raise RuntimeError(f"hctr_emb_dispatch failed rc={rc}")
        D._count()
        return
    for r in routes:
        dst = dst_bufs[r.dst_rank].reshape(-1)
        if r.kind == 0:
            s2 = torch.as_strided(src, (r.rows, r.row_elems), (r.src_stride, 1), src.storage_offset() + r.src_off)
            d2 = torch.as_strided(dst, (r.rows, r.row_elems), (r.dst_stride, 1), dst.storage_offset() + r.dst_off)
            d2.copy_(s2)
            continue
        keys = torch.as_strided(src, (r.rows, r.row_elems), (r.src_stride, 1),
                                src.storage_offset() + r.src_off).long()
        m = (keys >= 0) & (keys % r.k == r.shard)
        order = torch.argsort((~m).to(torch.int8), dim=1, stable=True)
        rows = torch.div(keys, r.k, rounding_mode="floor").gather(1, order)
        cnt = m.sum(1)
        pos = torch.arange(r.row_elems).view(1, -1)
        d2 = torch.as_strided(dst, (r.rows, r.row_elems), (r.dst_stride, 1), dst.storage_offset() + r.dst_off)
        d2.copy_(torch.where(pos < cnt.view(-1, 1), rows, torch.full_like(rows, -1)).to(dst.dtype))
        nn = nnz_bufs[r.dst_rank].reshape(-1)
        nn[r.nnz_off:r.nnz_off + r.rows].copy_(cnt.to(torch.int32))


# ----------------------------------------------------------------------------- forward
def forward(lookups, lookups_dev, table, ev_pitch, key_bufs, out_bufs, batch, my_rank=0,
            nnz_bufs=None, key_bytes=4, act_bf16=True):
    """key_bufs / out_bufs: one entry per source rank (tensor, or int device pointer for peers)."""
    R = len(key_bufs)
    if table.is_cuda:
        p = CEmbParams()
        p.num_ranks, p.my_rank, p.batch, p.num_lookups = R, my_rank, batch, len(lookups)
        kb = key_bytes
        out_bf16 = int(act_bf16)
        for r in range(R):
            k, o = key_bufs[r], out_bufs[r]
            p.keys[r] = k if isinstance(k, int) else k.data_ptr()
            p.out[r] = o if isinstance(o, int) else o.data_ptr()
            if nnz_bufs is not None:
                n = nnz_bufs[r]
                p.nnz[r] = n if isinstance(n, int) else n.data_ptr()
            if not isinstance(k, int):
                kb = k.element_size()
            if not isinstance(o, int):
                out_bf16 = int(o.dtype == torch.bfloat16)
        p.lookups = lookups_dev.data_ptr()
        p.table = table.data_ptr()
        p.ev_size = ev_pitch
        max_ev = max((l.ev_size for l in lookups), default=ev_pitch)
        rc = lib().hctr_emb_forward(C.byref(p), max_ev, kb, out_bf16, _st(table.device))
        if rc:
            raise RuntimeError("hctr_emb_forward failed")
        D._count()
        return
    forward_reference(lookups, table, ev_pitch, key_bufs, out_bufs, batch, nnz_bufs)


def _owned(lk: LookupDesc, keys):
    keys = keys.long()
    own = keys >= 0
    if lk.num_shards > 1:
        own &= (keys % lk.num_shards) == lk.shard_idx
    r = torch.div(keys, lk.num_shards, rounding_mode="floor")
    own &= r < lk.rows
    return own, lk.table_row_off + torch.where(own, r, torch.zeros_like(r))


def forward_reference(lookups, table, ev_pitch, key_bufs, out_bufs, batch, nnz_bufs=None):
    tab = table.view(-1, ev_pitch)
    for src in range(len(key_bufs)):
        kbuf, obuf = key_bufs[src].reshape(-1), out_bufs[src].reshape(-1)
        for lk in lookups:
            keys = torch.as_strided(kbuf, (batch, lk.hotness), (lk.key_stride, 1),
                                    kbuf.storage_offset() + lk.key_off)
            own, rows = _owned(lk, keys)
            nnz = torch.full((batch,), lk.hotness, device=keys.device)
            if lk.nnz_off >= 0 and nnz_bufs is not None:
                nnz = nnz_bufs[src].reshape(-1)[lk.nnz_off:lk.nnz_off + batch].long().clamp(max=lk.hotness)
                own &= torch.arange(lk.hotness, device=keys.device).view(1, -1) < nnz.view(-1, 1)
            v = tab[rows][..., :lk.ev_size].float() * own.unsqueeze(-1)
            pooled = v.sum(1)
            if lk.combiner == 1:
                pooled = pooled / nnz.clamp(min=1).view(-1, 1).float()
            o = torch.as_strided(obuf, (batch, lk.ev_size), (lk.out_stride, 1),
                                 obuf.storage_offset() + lk.out_off)
            o.copy_(pooled.to(o.dtype))


# ----------------------------------------------------------------------------- backward
class UniqueWorkspace:
    """Transient hash + fp32 per-unique-row gradient accumulator (device)."""

    def __init__(self, max_pairs: int, ev_pitch: int, device, unique_ratio: float = 1.0,
                 indexed: bool = False, need_scale: bool = False):
        self.device = device
        self.ev = ev_pitch
        self.max_unique = max(16, int(max_pairs * unique_ratio))
        cap = 1
        while cap < 2 * self.max_unique:
            cap <<= 1
        self.capacity = cap
        if device.type == "cuda":
            self.keys = torch.full((cap,), -1, dtype=torch.int64, device=device)
            self.vals = torch.full((cap,), -1, dtype=torch.int32, device=device)
            self.counter = torch.zeros(1, dtype=torch.int32, device=device)
            self.rows = torch.zeros(self.max_unique, dtype=torch.int64, device=device)
            self.slots = torch.zeros(self.max_unique, dtype=torch.int32, device=device)
            self.wgrad = None if indexed else torch.zeros(self.max_unique, ev_pitch,
                                                          dtype=torch.float32, device=device)
            self.overflow = torch.zeros(1, dtype=torch.int32, device=device)
            self.c = CUniqueTable(self.keys.data_ptr(), self.vals.data_ptr(),
                                  self.counter.data_ptr(), self.rows.data_ptr(),
                                  self.slots.data_ptr(), cap - 1, self.max_unique)
            self.indexed = indexed
            if indexed:
                # counting-sort index (csrc/embedding_bwd.cu); the fp32 per-row accumulator is
                # not needed on this path
                self.wgrad = None
                mu, mp = self.max_unique, max(int(max_pairs), 1)
                i32 = dict(dtype=torch.int32, device=device)
                self.count = torch.zeros(mu, **i32)
                self.offsets = torch.zeros(mu + 1, **i32)
                self.block_sums = torch.zeros((mu + 1023) // 1024 + 1, **i32)
                self.pair_uid = torch.full((mp,), -1, **i32)
                self.bucket_list = torch.zeros(mp, **i32)
                self.bucket_scale = torch.ones(mp, dtype=torch.float32, device=device) if need_scale else None
                self.max_heavy_rows = 8192
                self.max_heavy_items = 65536
                self.heavy_items = torch.zeros(2 * self.max_heavy_items, **i32)
                self.heavy_count = torch.zeros(1, **i32)
                self.heavy_slot = torch.zeros(1, **i32)
                self.heavy_scratch = torch.zeros(self.max_heavy_rows, ev_pitch, dtype=torch.float32,
                                                 device=device)
                self.heavy_ticket = torch.zeros(self.max_heavy_rows, **i32)
                self.light_list = torch.zeros(2 * mu, **i32)
                self.medium_list = torch.zeros(2 * mu, **i32)
                self.class_count = torch.zeros(2, **i32)
                self.ix = CBwdIndex(self.count.data_ptr(), self.offsets.data_ptr(),
                                    self.block_sums.data_ptr(), self.pair_uid.data_ptr(),
                                    self.bucket_list.data_ptr(),
                                    0 if self.bucket_scale is None else self.bucket_scale.data_ptr(),
                                    self.heavy_items.data_ptr(),
                                    self.heavy_count.data_ptr(), self.heavy_scratch.data_ptr(),
                                    self.heavy_ticket.data_ptr(), self.heavy_slot.data_ptr(),
                                    self.max_heavy_rows, self.max_heavy_items,
                                    self.light_list.data_ptr(), self.medium_list.data_ptr(),
                                    self.class_count.data_ptr())
        else:
            self.ref_rows = None
            self.ref_grads = None


def backward_accum(lookups, lookups_dev, table, ev_pitch, key_bufs, grad_bufs, batch, ws, grad_scale=1.0,
                   my_rank=0, nnz_bufs=None, dense_wgrad=None, key_bytes=4, act_bf16=True):
    """Accumulate bucket gradients per unique row (or into a dense wgrad for DP tables)."""
    R = len(key_bufs)
    if table.is_cuda:
        p = CEmbParams()
        p.num_ranks, p.my_rank, p.batch, p.num_lookups = R, my_rank, batch, len(lookups)
        kb, gbf = key_bytes, int(act_bf16)
        for r in range(R):
            k, g = key_bufs[r], grad_bufs[r]
            p.keys[r] = k if isinstance(k, int) else k.data_ptr()
            p.grad[r] = g if isinstance(g, int) else g.data_ptr()
            if nnz_bufs is not None:
                n = nnz_bufs[r]
                p.nnz[r] = n if isinstance(n, int) else n.data_ptr()
            if not isinstance(k, int):
                kb = k.element_size()
            if not isinstance(g, int):
                gbf = int(g.dtype == torch.bfloat16)
        p.lookups = lookups_dev.data_ptr()
        p.table = table.data_ptr()
        p.ev_size = ev_pitch
        max_ev = max((l.ev_size for l in lookups), default=ev_pitch)
        if dense_wgrad is not None:
            ut = CUniqueTable(0, 0, 0, 0, 0, 0, 0xFFFFFFFF)
            dst = dense_wgrad.data_ptr()
        else:
            ut, dst = ws.c, ws.wgrad.data_ptr()
        esz = 2 if gbf else 4
        ev4 = int(all(l.ev_size % 4 == 0 and (l.grad_off * esz) % 16 == 0 and
                      (l.grad_stride * esz) % 16 == 0 for l in lookups))
        rc = lib().hctr_emb_backward_accum(C.byref(p), C.byref(ut), dst, float(grad_scale), max_ev,
                                           kb, gbf, ev4, _st(table.device))
        if rc:
            raise RuntimeError("hctr_emb_backward_accum failed")
        D._count()
        return
    rows_all, grads_all = [], []
    for src in range(R):
        kbuf, gbuf = key_bufs[src].reshape(-1), grad_bufs[src].reshape(-1)
        for lk in lookups:
            keys = torch.as_strided(kbuf, (batch, lk.hotness), (lk.key_stride, 1),
                                    kbuf.storage_offset() + lk.key_off)
            own, rows = _owned(lk, keys)
            nnz = torch.full((batch,), lk.hotness, device=keys.device)
            if lk.nnz_off >= 0 and nnz_bufs is not None:
                nnz = nnz_bufs[src].reshape(-1)[lk.nnz_off:lk.nnz_off + batch].long().clamp(max=lk.hotness)
                own &= torch.arange(lk.hotness, device=keys.device).view(1, -1) < nnz.view(-1, 1)
            g = torch.as_strided(gbuf, (batch, lk.ev_size), (lk.grad_stride, 1),
                                 gbuf.storage_offset() + lk.grad_off)
            g = g.float() * grad_scale
            if lk.combiner == 1:
                g = g / nnz.clamp(min=1).view(-1, 1).float()
            gg = g.unsqueeze(1).expand(batch, lk.hotness, lk.ev_size)[own]
            if lk.ev_size < ev_pitch:
                gg = torch.nn.functional.pad(gg, (0, ev_pitch - lk.ev_size))
            rows_all.append(rows[own])
            grads_all.append(gg)
    rows = torch.cat(rows_all) if rows_all else torch.zeros(0, dtype=torch.long)
    grads = torch.cat(grads_all) if grads_all else torch.zeros(0, ev_pitch)
    if dense_wgrad is not None:
        dense_wgrad.view(-1, ev_pitch).index_add_(0, rows, grads)
        return
    if ws.ref_rows is not None:
        rows = torch.cat([ws.ref_rows, rows])
        grads = torch.cat([ws.ref_grads, grads])
    ws.ref_rows, ws.ref_grads = rows, grads


def sparse_opt_reference(opt: Optimizer_t, w, s0, s1, g, hp, lr, step):
    """Row-block update (w, s0, s1, g all [n, ev] fp32). Returns nothing, updates in place."""
    g = g / hp.get("scaler", 1.0)
    eps = hp.get("epsilon", 1e-7)
    if opt == Optimizer_t.SGD:
        w -= lr * g
    elif opt == Optimizer_t.AdaGrad:
        s0 += g * g
        w -= lr * g / (s0.sqrt() + eps)
    elif opt == Optimizer_t.Adam:
        b1, b2 = hp.get("beta1", 0.9), hp.get("beta2", 0.999)
        s0.mul_(b1).add_((1 - b1) * g)
        s1.mul_(b2).add_((1 - b2) * g * g)
        alpha = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
        w -= alpha * s0 / (s1.sqrt() + eps)
    elif opt == Optimizer_t.Ftrl:
        fb, l1, l2 = hp.get("ftrl_beta", 0.0), hp.get("lambda1", 0.0), hp.get("lambda2", 0.0)
        n_new = s1 + g * g
        s0 += g + ((s1 + fb).sqrt() - (n_new + fb).sqrt()) * w / lr
        s1.copy_(n_new)
        p = torch.where(s0 > 0, l1 - s0, -l1 - s0)
        q = (n_new + fb).sqrt() / lr + l2
        w.copy_(torch.where(s0.abs() > l1, p / q, torch.zeros_like(w)))
    elif opt == Optimizer_t.MomentumSGD:
        s0.mul_(hp.get("momentum", 0.0)).sub_(lr * g)
        w += s0
    elif opt == Optimizer_t.Nesterov:
        mu = hp.get("momentum", 0.0)
        an = mu * s0 - lr * g
        w += -mu * s0 + (1 + mu) * an
        s0.copy_(an)
    elif opt == Optimizer_t.RMSProp:
        b2 = hp.get("beta2", 0.999)
        s0.mul_(b2).add_((1 - b2) * g * g)
        w -= lr * g / (s0.sqrt() + eps)


def update(opt: Optimizer_t, table, s0, s1, ev_pitch, ws: UniqueWorkspace, hp: dict, lr_t, step_t,
           num_sms: int = 148):
    """Fused per-unique-row optimizer; consumes and clears the workspace."""
    if table.is_cuda:
        h = COptHyper(lr_t.data_ptr(), 1.0, hp.get("scaler", 1.0), hp.get("beta1", 0.9),
                      hp.get("beta2", 0.999), hp.get("epsilon", 1e-7), hp.get("lambda1", 0.0),
                      hp.get("lambda2", 0.0), hp.get("ftrl_beta", 0.0), hp.get("momentum", 0.0),
                      hp.get("initial_accu_value", 0.0), step_t.data_ptr())
        sbf = int(s0 is not None and s0.dtype == torch.bfloat16)
        rc = lib().hctr_emb_update(table.data_ptr(), 0 if s0 is None else s0.data_ptr(),
                                   0 if s1 is None else s1.data_ptr(), ws.wgrad.data_ptr(),
                                   C.byref(ws.c), ev_pitch, OPT_CODE[opt], sbf, C.byref(h),
                                   ws.overflow.data_ptr(), num_sms, _st(table.device))
        if rc:
            raise RuntimeError("hctr_emb_update failed")
        D._count(2)
        return
    if ws.ref_rows is None or ws.ref_rows.numel() == 0:
        ws.ref_rows = ws.ref_grads = None
        return
    update_rows(opt, table, s0, s1, ev_pitch, ws.ref_rows, ws.ref_grads, hp, lr_t, step_t)
    ws.ref_rows = ws.ref_grads = None


def update_rows(opt: Optimizer_t, table, s0, s1, ev_pitch, rows, grads, hp: dict, lr_t, step_t):
    """One optimizer application per DISTINCT row (``rows`` [n] may repeat, ``grads`` [n, pitch] fp32 are
    summed per row first).  Plain tensor ops on the tables' device: the CPU path of ``update`` and the
    owner side of the Unique-compression exchange (embedding/unique_exchange.py)."""
    if rows.numel() == 0:
        return
    uniq, inv = torch.unique(rows, return_inverse=True)
    g = torch.zeros(uniq.numel(), ev_pitch, dtype=torch.float32, device=grads.device).index_add_(0, inv, grads.float())
    tab = table.view(-1, ev_pitch)
    w = tab[uniq].clone()
    a = None if s0 is None else s0.view(-1, ev_pitch)[uniq].float().clone()
    b = None if s1 is None else s1.view(-1, ev_pitch)[uniq].float().clone()
    sparse_opt_reference(opt, w, a, b, g, hp, float(lr_t.item()), int(step_t.item()))
    tab[uniq] = w
    if a is not None:
        s0.view(-1, ev_pitch)[uniq] = a.to(s0.dtype)
    if b is not None:
        s1.view(-1, ev_pitch)[uniq] = b.to(s1.dtype)


def gather_rows(table, ev_pitch, rows, out):
    """out[i] = table[rows[i]] (rows < 0 -> zeros)."""
    if table.is_cuda and rows.dtype == torch.int64:
        rc = lib().hctr_emb_gather_rows(table.data_ptr(), rows.data_ptr(), out.data_ptr(),
                                        rows.numel(), ev_pitch, int(out.dtype == torch.bfloat16),
                                        _st(table.device))
        if rc:
            raise RuntimeError("hctr_emb_gather_rows failed")
        D._count()
        return
    v = table.view(-1, ev_pitch)[rows.clamp(min=0)] * (rows >= 0).unsqueeze(-1)
    out.copy_(v.to(out.dtype))


def _fill_params(lookups, lookups_dev, table, ev_pitch, key_bufs, grad_bufs, batch, my_rank, nnz_bufs,
                 key_bytes=4, act_bf16=True):
    R = len(key_bufs)
    p = CEmbParams()
    p.num_ranks, p.my_rank, p.batch, p.num_lookups = R, my_rank, batch, len(lookups)
    kb, gbf = key_bytes, int(act_bf16)
    for r in range(R):
        k = key_bufs[r]
        p.keys[r] = k if isinstance(k, int) else k.data_ptr()
        if not isinstance(k, int):
            kb = k.element_size()
        if grad_bufs is not None:
            g = grad_bufs[r]
            p.grad[r] = g if isinstance(g, int) else g.data_ptr()
            if not isinstance(g, int):
                gbf = int(g.dtype == torch.bfloat16)
        if nnz_bufs is not None:
            n = nnz_bufs[r]
            p.nnz[r] = n if isinstance(n, int) else n.data_ptr()
    p.lookups = lookups_dev.data_ptr()
    p.table = table.data_ptr()
    p.ev_size = ev_pitch
    return p, kb, gbf


def bwd_index(lookups, lookups_dev, table, ev_pitch, key_bufs, batch, ws, my_rank=0, nnz_bufs=None,
              key_bytes=4):
    """Gradient-independent part of the backward: unique rows + per-row bucket lists."""
    p, kb, _ = _fill_params(lookups, lookups_dev, table, ev_pitch, key_bufs, None, batch, my_rank,
                            nnz_bufs, key_bytes)
    total_pairs = len(key_bufs) * batch * sum(l.hotness for l in lookups)
    rc = lib().hctr_emb_bwd_index(C.byref(p), C.byref(ws.c), C.byref(ws.ix), total_pairs, kb,
                                  _st(table.device))
    if rc:
        raise RuntimeError("hctr_emb_bwd_index failed")
    D._count(6)


def bwd_reduce_update(opt: Optimizer_t, lookups, lookups_dev, table, s0, s1, ev_pitch, key_bufs,
                      grad_bufs, batch, ws, hp: dict, lr_t, step_t, grad_scale=1.0, my_rank=0,
                      nnz_bufs=None, num_sms: int = 148, key_bytes=4, act_bf16=True):
    """Fused gradient gather (local / peer loads) + duplicate reduction + optimizer."""
    p, kb, gbf = _fill_params(lookups, lookups_dev, table, ev_pitch, key_bufs, grad_bufs, batch,
                              my_rank, nnz_bufs, key_bytes, act_bf16)
    h = COptHyper(lr_t.data_ptr(), 1.0, hp.get("scaler", 1.0), hp.get("beta1", 0.9),
                  hp.get("beta2", 0.999), hp.get("epsilon", 1e-7), hp.get("lambda1", 0.0),
                  hp.get("lambda2", 0.0), hp.get("ftrl_beta", 0.0), hp.get("momentum", 0.0),
                  hp.get("initial_accu_value", 0.0), step_t.data_ptr())
    sbf = int(s0 is not None and s0.dtype == torch.bfloat16)
    rc = lib().hctr_emb_bwd_reduce_update(C.byref(p), C.byref(ws.c), C.byref(ws.ix),
                                          0 if s0 is None else s0.data_ptr(),
                                          0 if s1 is None else s1.data_ptr(), OPT_CODE[opt], sbf,
                                          C.byref(h), float(grad_scale), gbf, ws.overflow.data_ptr(),
                                          num_sms, _st(table.device))
    if rc:
        raise RuntimeError(f"hctr_emb_bwd_reduce_update failed rc={rc}")
    D._count(3)
