"""HashTable: key -> dense row index with insertion on first sight (legacy SparseEmbedding
semantics).  GPU: csrc/hashtable.cu (open addressing, atomicCAS); CPU: python dict reference.
Parity: HugeCTR/src/hashtable/nv_hashtable.cu:36-306 (get_insert / get(get_mark) / set / dump /
get_size), capacity = max_vocabulary_size / 0.75 (nv_hashtable.hpp:179).
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _native
from ..ops import dense as D

_lib = None


def lib():
    global _lib
    if _lib is None:
        l = _native.cuda_lib()
        vp, ull, ll, i = C.c_void_p, C.c_ulonglong, C.c_longlong, C.c_int
        l.hctr_ht_get_insert.argtypes = [vp, vp, vp, ull, ll, vp, vp, ll, i, vp, vp]
        l.hctr_ht_set.argtypes = [vp, vp, vp, ull, vp, vp, ll, vp]
        l.hctr_ht_dump.argtypes = [vp, vp, ull, vp, vp, vp, vp]
        l.hctr_ht_translate.argtypes = [vp, vp, vp, ull, ll, vp, vp, ll, i, vp, ll, i, i, i, i, vp]
        l.hctr_ht_translate.restype = i
        for n in ("hctr_ht_get_insert", "hctr_ht_set", "hctr_ht_dump"):
            getattr(l, n).restype = i
        _lib = l
    return _lib


class HashTable:
    LOAD_FACTOR = 0.75

    def __init__(self, max_rows: int, device):
        self.max_rows = int(max_rows)
        self.device = torch.device(device)
        cap = 1
        while cap < max(16, int(self.max_rows / self.LOAD_FACTOR)):
            cap <<= 1
        self.capacity = cap
        if self.device.type == "cuda":
            self.keys = torch.full((cap,), -1, dtype=torch.int64, device=self.device)
            self.vals = torch.full((cap,), -1, dtype=torch.int64, device=self.device)
            self.counter = torch.zeros(1, dtype=torch.int64, device=self.device)
            self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)   # sticky device flag
        else:
            self.map = {}
            self._overflowed = False

    def _st(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def get_insert(self, keys: torch.Tensor) -> torch.Tensor:
        """rows for keys; unseen keys get the next free row. keys < 0 -> -1. Overflow -> -1 and
        size() > max_rows (check_overflow raises)."""
        return self._lookup(keys, True)

    def get(self, keys: torch.Tensor) -> torch.Tensor:
        """lookup only (get_mark): missing keys -> -1 (treated as zero vectors by the lookup)."""
        return self._lookup(keys, False)

    def _lookup(self, keys, insert):
        k = keys.reshape(-1).to(torch.int64)
        if self.device.type == "cuda":
            out = torch.empty_like(k)
            rc = lib().hctr_ht_get_insert(self.keys.data_ptr(), self.vals.data_ptr(),
                                          self.counter.data_ptr(), self.capacity, self.max_rows,
                                          k.data_ptr(), out.data_ptr(), k.numel(), int(insert),
                                          self._st(), self.overflow.data_ptr())
            if rc:
                raise RuntimeError("hash get_insert failed")
            D._count()
            return out.view(keys.shape)
        out = torch.empty_like(k)
        kl = k.tolist()
        res = []
        for x in kl:
            if x < 0:
                res.append(-1)
            elif x in self.map:
                v = self.map[x]
                res.append(v if v < self.max_rows else -1)
            elif insert:
                v = len(self.map)
                if v >= self.max_rows:          # rows exhausted: no insertion, sticky overflow
                    self._overflowed = True
                    res.append(-1)
                    continue
                self.map[x] = v
                res.append(v)
            else:
                res.append(-1)
        return torch.tensor(res, dtype=torch.int64).view(keys.shape)

    def translate(self, keys: torch.Tensor, out: torch.Tensor, insert: bool, n_per_rank: int, key_mod: int = 1,
                  key_rem: int = 0, slot_div: int = 1, slot_num: int = 0):
        """out = row of every OWNED key of ``keys`` (int64, any shape; -1 elsewhere) in one launch, no host
        sync.  Ownership: ``key % key_mod == key_rem`` or, with ``slot_num`` > 0, the slot of the key's
        position ``((pos % n_per_rank) // slot_div) % slot_num) % key_mod == key_rem``."""
        k = keys.reshape(-1)
        if self.device.type == "cuda":
            rc = lib().hctr_ht_translate(self.keys.data_ptr(), self.vals.data_ptr(), self.counter.data_ptr(),
                                         self.capacity, self.max_rows, k.data_ptr(), out.data_ptr(), k.numel(),
                                         int(insert), self.overflow.data_ptr(), int(n_per_rank), int(key_mod),
                                         int(key_rem), int(slot_div), int(slot_num), self._st())
            if rc:
                raise RuntimeError("hash translate failed")
            D._count()
            return out
        pos = torch.arange(k.numel()) % max(int(n_per_rank), 1)
        own = k >= 0
        if slot_num > 0:
            own &= ((pos // slot_div) % slot_num) % key_mod == key_rem
        elif key_mod > 1:
            own &= (k % key_mod) == key_rem
        kk = torch.where(own, k, torch.full_like(k, -1))
        out.reshape(-1).copy_(self._lookup(kk, insert))
        return out

    def set(self, keys: torch.Tensor, vals: torch.Tensor):
        k, v = keys.reshape(-1).to(torch.int64), vals.reshape(-1).to(torch.int64)
        if self.device.type == "cuda":
            k, v = k.to(self.device), v.to(self.device)
            rc = lib().hctr_ht_set(self.keys.data_ptr(), self.vals.data_ptr(), self.counter.data_ptr(),
                                   self.capacity, k.data_ptr(), v.data_ptr(), k.numel(), self._st())
            if rc:
                raise RuntimeError("hash set failed")
            self.counter.fill_(max(int(self.counter.item()), int(v.max().item()) + 1 if v.numel() else 0))
        else:
            for a, b in zip(k.tolist(), v.tolist()):
                self.map[a] = b

    def overflowed(self) -> bool:
        """True once a key could not get a row (host sync on CUDA: call it outside the step)."""
        if self.device.type == "cuda":
            return bool(int(self.overflow.item()))
        return self._overflowed

    def size(self) -> int:
        if self.device.type == "cuda":
            return int(self.counter.item())
        return len(self.map)

    def dump(self):
        """-> (keys int64 [n], rows int64 [n]) on CPU"""
        if self.device.type == "cuda":
            n = self.capacity
            ok = torch.empty(n, dtype=torch.int64, device=self.device)
            ov = torch.empty(n, dtype=torch.int64, device=self.device)
            cnt = torch.zeros(1, dtype=torch.int64, device=self.device)
            lib().hctr_ht_dump(self.keys.data_ptr(), self.vals.data_ptr(), self.capacity,
                               ok.data_ptr(), ov.data_ptr(), cnt.data_ptr(), self._st())
            c = int(cnt.item())
            k, v = ok[:c].cpu(), ov[:c].cpu()
        else:
            k = torch.tensor(list(self.map.keys()), dtype=torch.int64)
            v = torch.tensor(list(self.map.values()), dtype=torch.int64)
        order = torch.argsort(v)
        return k[order], v[order]

    def clear(self):
        if self.device.type == "cuda":
            self.keys.fill_(-1)
            self.vals.fill_(-1)
            self.counter.zero_()
            self.overflow.zero_()
        else:
            self.map = {}
            self._overflowed = False
