"""gpu_cache: set-associative LRU cache of embedding rows in HBM + StaticTable + UvmTable.

Parity: gpu_cache/include/nv_gpu_cache.hpp (Query / Replace / Update / Dump),
gpu_cache/src/static_hash_table.cu:125-495 (read-only table), gpu_cache/src/uvm_table.cu:39-606
(HBM hot tier + pinned-host cold tier).  GPU kernels: csrc/gpu_cache.cu; the CPU implementation
below has identical semantics (it is the test oracle).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch

from .. import _native
from ..ops import dense as D

_lib = None


def lib():
    global _lib
    if _lib is None:
        l = _native.cuda_lib()
        vp, i, ll = C.c_void_p, C.c_int, C.c_longlong
        common = [vp, vp, vp, vp, vp, i, i, i]
        l.hctr_cache_query.argtypes = common + [vp, ll, vp, vp, vp, vp, vp]
        l.hctr_cache_replace.argtypes = common + [vp, vp, ll, vp, vp, vp, vp]
        l.hctr_cache_update.argtypes = common + [vp, vp, ll, vp]
        l.hctr_cache_dump.argtypes = common + [i, i, vp, vp, vp]
        for n in ("hctr_cache_query", "hctr_cache_replace", "hctr_cache_update", "hctr_cache_dump"):
            getattr(l, n).restype = i
        _lib = l
    return _lib


class GpuCache:
    """capacity = num_sets * ways rows of ``ev`` floats; ways is a multiple of 32."""

    def __init__(self, capacity_rows: int, ev: int, device, ways: int = 64):
        self.device = torch.device(device)
        self.ev = int(ev)
        self.ways = max(32, (ways + 31) // 32 * 32)
        self.num_sets = max(1, (int(capacity_rows) + self.ways - 1) // self.ways)
        n = self.num_sets * self.ways
        self.capacity = n
        if self.device.type == "cuda":
            dev = self.device
            self.keys = torch.full((n,), -1, dtype=torch.int64, device=dev)
            self.stamps = torch.zeros(n, dtype=torch.int64, device=dev)
            self.vals = torch.zeros(n, ev, dtype=torch.float32, device=dev)
            self.locks = torch.zeros(self.num_sets, dtype=torch.int32, device=dev)
            self.clock = torch.zeros(1, dtype=torch.int64, device=dev)
            self._cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        else:
            self.sets = [OrderedDict() for _ in range(self.num_sets)]
        self.hits = 0
        self.queries = 0

    def _args(self):
        return (self.keys.data_ptr(), self.stamps.data_ptr(), self.vals.data_ptr(),
                self.locks.data_ptr(), self.clock.data_ptr(), self.num_sets, self.ways, self.ev)

    def _st(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    @staticmethod
    def _hash(k: int) -> int:
        k &= 0xFFFFFFFFFFFFFFFF
        k ^= k >> 33
        k = (k * 0xff51afd7ed558ccd) & 0xFFFFFFFFFFFFFFFF
        k ^= k >> 33
        k = (k * 0xc4ceb9fe1a85ec53) & 0xFFFFFFFFFFFFFFFF
        k ^= k >> 33
        return k & 0xFFFFFFFF

    # ------------------------------------------------------------------ API
    def query(self, keys: torch.Tensor):
        """-> (values [n, ev] (hit rows filled), missing_index [m], missing_keys [m])"""
        k = keys.reshape(-1).to(torch.int64)
        n = k.numel()
        self.queries += n
        if self.device.type == "cuda":
            out = torch.zeros(n, self.ev, device=self.device)
            mi = torch.empty(n, dtype=torch.int64, device=self.device)
            mk = torch.empty(n, dtype=torch.int64, device=self.device)
            self._cnt.zero_()
            rc = lib().hctr_cache_query(*self._args(), k.data_ptr(), n, out.data_ptr(), mi.data_ptr(),
                                        mk.data_ptr(), self._cnt.data_ptr(), self._st())
            if rc:
                raise RuntimeError("cache query failed")
            D._count()
            m = int(self._cnt.item())
            self.hits += n - m
            order = torch.argsort(mi[:m])
            return out, mi[:m][order], mk[:m][order]
        out = torch.zeros(n, self.ev)
        mi, mk = [], []
        for i, x in enumerate(k.tolist()):
            s = self.sets[self._hash(x) % self.num_sets]
            if x >= 0 and x in s:
                s.move_to_end(x)
                out[i] = s[x]
                self.hits += 1
            else:
                mi.append(i)
                mk.append(x)
        return out, torch.tensor(mi, dtype=torch.int64), torch.tensor(mk, dtype=torch.int64)

    def replace(self, keys: torch.Tensor, values: torch.Tensor, return_evicted: bool = False):
        k = keys.reshape(-1).to(torch.int64)
        v = values.reshape(-1, self.ev).to(torch.float32)
        n = k.numel()
        if self.device.type == "cuda":
            ek = ev = None
            if return_evicted:
                ek = torch.empty(n, dtype=torch.int64, device=self.device)
                ev = torch.empty(n, self.ev, device=self.device)
            self._cnt.zero_()
            rc = lib().hctr_cache_replace(*self._args(), k.data_ptr(), v.contiguous().data_ptr(), n,
                                          0 if ek is None else ek.data_ptr(),
                                          0 if ev is None else ev.data_ptr(), self._cnt.data_ptr(),
                                          self._st())
            if rc:
                raise RuntimeError("cache replace failed")
            D._count()
            if return_evicted:
                m = int(self._cnt.item())
                return ek[:m], ev[:m]
            return None
        eks, evs = [], []
        for x, row in zip(k.tolist(), v):
            if x < 0:
                continue
            s = self.sets[self._hash(x) % self.num_sets]
            if x in s:
                s[x] = row.clone()
                s.move_to_end(x)
                continue
            if len(s) >= self.ways:
                ok, ov = s.popitem(last=False)
                eks.append(ok)
                evs.append(ov)
            s[x] = row.clone()
        if return_evicted:
            return (torch.tensor(eks, dtype=torch.int64),
                    torch.stack(evs) if evs else torch.zeros(0, self.ev))
        return None

    def update(self, keys: torch.Tensor, values: torch.Tensor):
        """overwrite rows that are present; absent keys are ignored"""
        k = keys.reshape(-1).to(torch.int64)
        v = values.reshape(-1, self.ev).to(torch.float32)
        if self.device.type == "cuda":
            rc = lib().hctr_cache_update(*self._args(), k.data_ptr(), v.contiguous().data_ptr(),
                                         k.numel(), self._st())
            if rc:
                raise RuntimeError("cache update failed")
            D._count()
            return
        for x, row in zip(k.tolist(), v):
            s = self.sets[self._hash(x) % self.num_sets]
            if x in s:
                s[x] = row.clone()

    def dump(self, start_set: int = 0, end_set: int = None) -> torch.Tensor:
        end_set = self.num_sets if end_set is None else end_set
        if self.device.type == "cuda":
            n = (end_set - start_set) * self.ways
            out = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
            self._cnt.zero_()
            lib().hctr_cache_dump(*self._args(), start_set, end_set, out.data_ptr(),
                                  self._cnt.data_ptr(), self._st())
            return out[:int(self._cnt.item())]
        ks = []
        for s in self.sets[start_set:end_set]:
            ks.extend(s.keys())
        return torch.tensor(ks, dtype=torch.int64)

    def hit_rate(self) -> float:
        return self.hits / max(1, self.queries)


class StaticTable:
    """Read-only key->vector table fully resident in HBM ("> 3x faster than the embedding cache",
    release_notes.md:300): hash to dense rows once at build time, lookups are plain gathers."""

    def __init__(self, keys: torch.Tensor, values: torch.Tensor, device):
        from ..embedding.hashtable import HashTable
        self.device = torch.device(device)
        self.ev = values.shape[1]
        self.ht = HashTable(max(16, keys.numel()), self.device)
        rows = self.ht.get_insert(keys.to(self.device).to(torch.int64))
        self.table = torch.zeros(keys.numel(), self.ev, device=self.device)
        self.table[rows] = values.to(self.device, torch.float32)

    def lookup(self, keys: torch.Tensor, default: float = 0.0) -> torch.Tensor:
        rows = self.ht.get(keys.to(self.device).to(torch.int64)).reshape(-1)
        out = self.table[rows.clamp(min=0)]
        out[rows < 0] = default
        return out


class UvmTable:
    """Two tiers: the ``hbm_rows`` most frequent keys live in HBM (StaticTable), the rest in pinned
    host memory and are fetched on demand (gpu_cache/src/uvm_table.cu: hot HBM + cold host)."""

    def __init__(self, keys: torch.Tensor, values: torch.Tensor, device, hbm_rows: int,
                 frequencies: torch.Tensor = None):
        n = keys.numel()
        order = torch.argsort(frequencies, descending=True) if frequencies is not None else torch.arange(n)
        hot, cold = order[:hbm_rows], order[hbm_rows:]
        self.device = torch.device(device)
        self.ev = values.shape[1]
        self.hot = StaticTable(keys[hot], values[hot], device) if hot.numel() else None
        self.cold_index = {int(k): i for i, k in enumerate(keys[cold].tolist())}
        cv = values[cold].to(torch.float32).contiguous()
        self.cold_vals = cv.pin_memory() if torch.cuda.is_available() else cv

    def lookup(self, keys: torch.Tensor) -> torch.Tensor:
        k = keys.reshape(-1).to(torch.int64)
        out = torch.zeros(k.numel(), self.ev, device=self.device)
        miss = torch.ones(k.numel(), dtype=torch.bool, device=self.device)
        if self.hot is not None:
            rows = self.hot.ht.get(k.to(self.device)).reshape(-1)
            hit = rows >= 0
            out[hit] = self.hot.table[rows[hit]]
            miss = ~hit
        mi = torch.nonzero(miss).reshape(-1).cpu()
        if mi.numel():
            kk = k.cpu()[mi].tolist()
            idx = torch.tensor([self.cold_index.get(x, -1) for x in kk], dtype=torch.int64)
            ok = idx >= 0
            if bool(ok.any()):
                vals = self.cold_vals[idx[ok]]
                out[mi[ok].to(self.device)] = vals.to(self.device, non_blocking=True)
        return out
