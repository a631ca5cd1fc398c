"""Legacy SparseEmbedding whose table lives on the HOST parameter server with the hot rows in the HBM
gpu_cache -- the Embedding Training Cache of the reference (``hugectr.CreateETC``, ``TrainPSType_t``:
HugeCTR/include/embedding_training_cache/{embedding_training_cache,parameter_server,hmem_cache}.hpp, headers
only in v25.03; semantics from the 22.x releases) rebuilt on the gpu_cache library the reference still ships
(gpu_cache/src/nv_gpu_cache.cu:154-1645: Query / Replace / Update) plus the native host parameter server
(csrc/host/param_server.cpp).

Per step, on the rank that OWNS a key (Distributed: key % N; Localized: slot % N):
  forward   unique owned keys -> cache.Query -> misses pulled from the host server (created on first sight) ->
            cache.Replace -> rows staged in a dense device block; the usual pool / exchange kernels run on it
  backward  fused reduce + optimizer on the staged rows (same kernels as the in-HBM tables) -> cache.Update +
            write-through of weights and optimizer states to the host server
Device memory = workspace_size_per_gpu_in_mb (cache + staging), table size is bounded by HOST memory only.
``TrainPSType_t.Staged`` stages without the cache tier (every step pulls / pushes its rows).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ..io import filesystem as FS

from ..enums import Optimizer_t, TrainPSType_t
from . import ops as E
from .sparse_embedding import SparseEmbeddingRuntime


class CachedSparseEmbeddingRuntime(SparseEmbeddingRuntime):
    graph_safe = False          # misses are served by the host: the step cannot be captured

    def __init__(self, cfg, param, layout, batch_per_gpu, device, act_dtype, comm, opt, key_dtype, scaler=1.0,
                 seed=0, share_from=None, ps_type=TrainPSType_t.Cached, host_capacity=0, local_path=None):
        super().__init__(cfg, param, layout, batch_per_gpu, device, act_dtype, comm, opt, key_dtype, scaler, seed,
                         share_from)
        self.ps_type = ps_type
        if share_from is not None:
            for k in ("ps", "cache", "state_cache", "nstates"):
                setattr(self, k, getattr(share_from, k))
            return
        from ..cache.gpu_cache import GpuCache
        from ..cache.hps import HostParameterServer
        self.nstates = sum(s is not None for s in (self.s0, self.s1))
        vocab = sum(cfg.slot_size_array) if cfg.slot_size_array else 0
        cap = int(host_capacity or max(vocab // max(1, self.world) + 1024, 4 * self.max_rows))
        bound = float(np.sqrt(1.0 / max(1.0, vocab / max(1, len(cfg.slot_size_array) or 1)))) if vocab else 0.05
        self.ps = HostParameterServer(self.vec, self.nstates, init_bound=bound, ssd_path=local_path,
                                      capacity_rows=cap, seed=seed * 7919 + 13)      # (rank independent: first-sight vectors are keyed by key)
        if self.nstates and self.opt.optimizer_type == Optimizer_t.AdaGrad and self.opt.initial_accu_value:
            self.ps.s[0].fill_(self.opt.initial_accu_value)
        # half of the workspace rows cache hot rows, the other half (self.table of the base class) stages
        # the rows of the current step
        cache_rows = max(64, self.max_rows // 2) if ps_type == TrainPSType_t.Cached else 0
        self.cache = GpuCache(cache_rows, self.vec, device) if cache_rows else None
        self.state_cache = [GpuCache(cache_rows, self.vec, device) for _ in range(self.nstates)] \
            if cache_rows else []

    def eval_clone(self, batch_per_gpu):
        return CachedSparseEmbeddingRuntime(self.cfg, self.param, self.layout, batch_per_gpu, self.device,
                                            self.act_dtype, self.comm, self.opt, self.key_dtype, self.scaler, 0,
                                            share_from=self, ps_type=self.ps_type)

    # ------------------------------------------------------------------ staging
    def _fetch(self, uniq: torch.Tensor, create: bool):
        """rows of ``uniq`` (device int64) -> (weights [U, vec], [states]) on the device"""
        dev = self.device
        U = uniq.numel()
        if self.cache is not None:
            vals, mi, mk = self.cache.query(uniq)
            sts = [c.query(uniq)[0] for c in self.state_cache]
        else:
            vals = torch.zeros(U, self.vec, device=dev)
            sts = [torch.zeros(U, self.vec, device=dev) for _ in range(self.nstates)]
            mi, mk = torch.arange(U, device=dev), uniq
        if mk.numel():
            mk_h = mk.cpu()
            rows = self.ps._rows(mk_h, create=create)
            known = (rows >= 0)
            # rows < 0 (unknown key in evaluation) gather as zeros; staging buffers are reused step after step
            vals[mi] = self.ps._gather(self.ps.w, rows, reuse=0).to(dev)
            for j, (t, src) in enumerate(zip(sts, self.ps.s)):
                t[mi] = self.ps._gather(src, rows, reuse=1 + j).to(dev)
            if self.cache is not None and bool(known.any()):
                kk = mk[known.to(dev)]
                self.cache.replace(kk, vals[mi][known.to(dev)])
                for c, t in zip(self.state_cache, sts):
                    c.replace(kk, t[mi][known.to(dev)])
        return vals, sts

    def forward(self, is_train: bool):
        W, b, S, H, vec = self.world, self.b, self.S, self.H, self.vec
        self.comm.all_gather(self.keys_all, self.keys_loc)
        own = self._owner_mask(self.keys_all)
        flat = self.keys_all.reshape(-1)
        ownf = own.reshape(-1)
        rows = torch.full_like(flat, -1)
        uniq, inv = torch.unique(flat[ownf], return_inverse=True)          # (host sync: sizes)
        U = uniq.numel()
        if U > self.max_rows:
            raise RuntimeError(f"{self.name}: {U} distinct keys in one step exceed the staging capacity "
                               f"({self.max_rows} rows): raise workspace_size_per_gpu_in_mb")
        if U:
            vals, sts = self._fetch(uniq, create=is_train)
            self.table.view(-1, vec)[:U] = vals
            for dst, src in zip((self.s0, self.s1), sts):
                if dst is not None:
                    dst.view(-1, vec)[:U] = src
            rows[ownf] = inv
        self._uniq = uniq
        self.rows_all.copy_(rows.view_as(self.rows_all))
        self.partial.zero_()
        E.forward(self.lookups, self.lookups_dev, self.table, vec, list(self.rows_all.unbind(0)),
                  list(self.partial.unbind(0)), b, self.rank)
        self.comm.all_to_all(self.recv, self.partial)
        out = self.recv.float().sum(0).view(b, S, vec)
        if self.combiner == 1:
            cnt = (self.keys_loc.view(b, S, H) >= 0).sum(-1).clamp(min=1).float()
            self._nnz_cnt = cnt
            out = out / cnt.unsqueeze(-1)
        self.top_data.copy_(out.to(self.top_data.dtype))

    def backward(self, lr_t, step_t):
        super().backward(lr_t, step_t)          # fused reduce + optimizer on the staged rows
        uniq = self._uniq
        U = uniq.numel()
        if not U:
            return
        vec = self.vec
        w = self.table.view(-1, vec)[:U]
        sts = [s.view(-1, vec)[:U] for s in (self.s0, self.s1) if s is not None]
        if self.cache is not None:
            self.cache.update(uniq, w)
            for c, s in zip(self.state_cache, sts):
                c.update(uniq, s)
        self.ps.push(uniq, w, sts)               # write-through

    def _global_sweep(self, lr_t, step_t):
        """(dense-equivalent Global updates would have to walk the host table: the offloaded table uses
        the Local rule -- only touched rows move -- like the reference's ETC)"""

    def check_overflow(self):
        pass

    # ------------------------------------------------------------------ checkpoint: the host server is the table
    def _gather_all(self):
        keys, rows = self.ps.items()
        order = torch.argsort(keys)
        keys, rows = keys[order], rows[order]
        w = self.ps.w[rows].clone()
        return self.comm.all_gather_object((keys, w, self._slot_ids(keys) if self.localized else None))

    def load_parameters(self, path: str):
        keys = torch.from_numpy(FS.read_array(FS.path_join(path, "key"), "<i8").astype("int64"))
        w = torch.from_numpy(FS.read_array(FS.path_join(path, "emb_vector"), "<f4")).view(-1, self.vec)
        if self.localized and FS.path_exists(FS.path_join(path, "slot_id")):
            slot = torch.from_numpy(FS.read_array(FS.path_join(path, "slot_id"), "<u8").astype("int64"))
            m = (slot % self.world) == self.rank
        else:
            m = (keys % self.world) == self.rank
        self._loaded_keys = keys
        self.ps.push(keys[m], w[m])
        if self.cache is not None:       # cached copies of reloaded rows are stale
            self.cache.keys.fill_(-1) if self.device.type == "cuda" else [s.clear() for s in self.cache.sets]

    def dump_opt_states(self, path: str):
        keys, rows = self.ps.items()
        order = torch.argsort(keys)
        rows = rows[order]
        st = [s[rows].clone() for s in self.ps.s]
        parts = self.comm.all_gather_object(st)
        if self.comm.rank == 0:
            FS.write_array(path, np.concatenate([torch.cat([p[i] for p in parts]).numpy().astype("<f4").reshape(-1)
                                                 for i in range(len(st))]) if st else np.zeros(0, "<f4"))
        self.comm.barrier()

    def load_opt_states(self, path: str):
        raw = FS.read_array(path, "<f4")
        if not self.ps.s:
            return
        keys = getattr(self, "_loaded_keys", None)
        if keys is None:
            parts = self.comm.all_gather_object(torch.sort(self.ps.items()[0]).values)
            keys = torch.cat(parts)
        tot = keys.numel()
        m = (keys % self.world) == self.rank if not self.localized else None
        if m is None:
            mine = set(self.ps.items()[0].tolist())
            m = torch.tensor([k in mine for k in keys.tolist()], dtype=torch.bool)
        per = tot * self.vec
        rows = self.ps._rows(keys[m], create=True)
        for i, s in enumerate(self.ps.s):
            blk = torch.from_numpy(raw[i * per:(i + 1) * per].copy()).view(tot, self.vec)
            self.ps._scatter(s, rows, blk[m])
