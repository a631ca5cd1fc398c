"""Norm-format reader with optional per-record checksum verification (Check_t.Sum).

File: DataSetHeader {error_check, number_of_records, label_dim, dense_dim, slot_num, reserved[3]}
(8 x int64, HugeCTR/include/common.hpp:184-191) followed by records
[label f32 x L][dense f32 x D] then per slot {int nnz, keys[nnz]}; with check_sum every record is
wrapped as {int nbytes, payload, char sum} (HugeCTR/include/data_generator.hpp:137-188,
include/data_readers/check_sum.hpp).  The reader side of component C46.

``NormReader`` is the native reader (csrc/host/norm_reader.cpp): a producer thread scans, verifies and
decodes this rank's slice of every global batch into a ring of pinned slots ahead of the consumer.
``PyNormReader`` is the pure-Python decoder of the same format, kept as the test oracle
(``HCTR_NORM_READER=python`` selects it).
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np
import torch

from ..enums import Check_t
from .batch import HostBatch
from .. import _native
from .parquet_reader import read_file_list
from .readers import IDataReader


class DataCheckError(RuntimeError):
    pass


class PyNormReader(IDataReader):
    def __init__(self, model, is_train: bool):
        rp = model.reader_params
        b = model.b_train if is_train else model.b_eval
        super().__init__(b, model.comm.rank, model.world, repeat=model.solver.repeat_dataset)
        self.file_list = rp.source[0] if is_train else rp.eval_source
        self.check = rp.check_type
        self.inp = model.input
        self.layout = model.layout
        self.key_dtype = model.key_dtype
        self.key_np = np.dtype("<i8") if model.solver.i64_input_key else np.dtype("<u4")
        self.slot_offsets = None
        n_slots = model.layout.total_slots
        if rp.slot_size_array and len(rp.slot_size_array) >= n_slots and model.sparse_embeddings:
            self.slot_offsets = np.concatenate([[0], np.cumsum(rp.slot_size_array[:n_slots])[:-1]]).astype("int64")
        self._it = None

    def _records(self):
        files = read_file_list(self.file_list)
        while True:
            for fp in files:
                with open(fp, "rb") as f:
                    raw = f.read()
                hdr = struct.unpack("<8q", raw[:64])
                chk, n, L, Dd, S = hdr[0], hdr[1], hdr[2], hdr[3], hdr[4]
                if L != self.inp.label_dim or Dd != self.inp.dense_dim or S != self.layout.total_slots:
                    raise RuntimeError(f"{fp}: header (label {L}, dense {Dd}, slots {S}) does not match the model")
                if chk == 1 and self.check != Check_t.Sum:
                    raise RuntimeError(f"{fp} was written with check_sum, use Check_t.Sum")
                pos = 64
                ks = self.key_np.itemsize
                for _ in range(n):
                    if chk == 1:
                        nb = struct.unpack_from("<i", raw, pos)[0]
                        payload = raw[pos + 4:pos + 4 + nb]
                        cs = struct.unpack_from("<b", raw, pos + 4 + nb)[0]
                        if np.frombuffer(payload, dtype=np.int8).sum(dtype=np.int8) != cs:
                            raise DataCheckError(f"{fp}: checksum mismatch")
                        pos += 5 + nb
                        rec, rp_ = payload, 0
                    else:
                        rec, rp_ = raw, pos
                    lab = np.frombuffer(rec, "<f4", L, rp_); rp_ += 4 * L
                    den = np.frombuffer(rec, "<f4", Dd, rp_); rp_ += 4 * Dd
                    slots = []
                    for s in range(S):
                        c = struct.unpack_from("<i", rec, rp_)[0]; rp_ += 4
                        slots.append(np.frombuffer(rec, self.key_np, c, rp_).astype("int64")); rp_ += ks * c
                    if chk != 1:
                        pos = rp_
                    yield lab, den, slots
            if not self.repeat:
                return

    def start(self):
        self._it = self._records()
        self.started = True

    def set_source(self, source=None):
        if source:
            self.file_list = source if isinstance(source, str) else source[0]
        self.start()

    def read_a_batch(self):
        if self._it is None:
            self.start()
        gb, b, r = self.b * self.world, self.b, self.rank
        recs = []
        for _ in range(gb):
            try:
                recs.append(next(self._it))
            except StopIteration:
                break
        if not recs:
            return None
        self.current_batchsize = len(recs)
        mine = recs[r * b:(r + 1) * b]
        L = torch.zeros(b, self.inp.label_dim)
        D = torch.zeros(b, self.inp.dense_dim)
        blocks, nnzs = [], []
        for i, (lab, den, _) in enumerate(mine):
            L[i] = torch.from_numpy(lab.copy())
            D[i] = torch.from_numpy(den.copy())
        si = 0
        for (name, S, H, fixed) in self.layout.blocks:
            blk = np.full((b, S, H), -1, dtype="int64")
            nz = np.zeros((S, b), dtype="int32")
            for s in range(S):
                off = 0 if self.slot_offsets is None else self.slot_offsets[si]
                for i, (_, _, slots) in enumerate(mine):
                    k = slots[si][:H]
                    blk[i, s, :len(k)] = k + off
                    nz[s, i] = len(k)
                si += 1
            blocks.append(torch.from_numpy(blk.reshape(-1)))
            nnzs.append(torch.from_numpy(nz.reshape(-1)))
        keys = torch.cat(blocks).to(self.key_dtype)
        return HostBatch(L, D, keys, torch.cat(nnzs), len(mine)).pin()



class NormReader(PyNormReader):
    """Native Norm reader; same constructor / batch layout as ``PyNormReader``."""

    DEPTH = 4

    def __new__(cls, model, is_train: bool):
        if os.environ.get("HCTR_NORM_READER", "native") == "python":
            return PyNormReader(model, is_train)
        return super().__new__(cls)

    def __init__(self, model, is_train: bool):
        super().__init__(model, is_train)
        self.h = None
        self.lib = L = _native.host_lib()
        vp, ip = C.c_void_p, C.POINTER(C.c_int)
        L.hctr_norm_open.restype = vp
        L.hctr_norm_open.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, ip, ip, C.c_int,
                                     C.POINTER(C.c_longlong), C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int]
        L.hctr_norm_start.argtypes = [vp, C.c_int] + [C.POINTER(vp)] * 4
        L.hctr_norm_next.argtypes = [vp, ip]
        L.hctr_norm_error.restype = C.c_char_p
        L.hctr_norm_error.argtypes = [vp]
        L.hctr_norm_close.argtypes = [vp]
        b = self.b
        nk = sum(b * S * H for (_, S, H, _) in self.layout.blocks)
        nn = sum(b * S for (_, S, H, _) in self.layout.blocks)
        pin = torch.cuda.is_available()
        self.slots = []
        for _ in range(self.DEPTH):
            ts = [torch.zeros(b, self.inp.label_dim), torch.zeros(b, self.inp.dense_dim),
                  torch.zeros(max(nk, 1), dtype=self.key_dtype), torch.zeros(max(nn, 1), dtype=torch.int32)]
            self.slots.append([t.pin_memory() if pin else t for t in ts])
        self.nk, self.nn = nk, nn

    def start(self):
        self.stop()
        files = read_file_list(self.file_list)
        arr = (C.c_char_p * len(files))(*[f.encode() for f in files])
        blocks = self.layout.blocks
        bs = (C.c_int * len(blocks))(*[S for (_, S, H, _) in blocks])
        bh = (C.c_int * len(blocks))(*[H for (_, S, H, _) in blocks])
        so = None
        if self.slot_offsets is not None:
            so = (C.c_longlong * len(self.slot_offsets))(*[int(v) for v in self.slot_offsets])
        self.h = self.lib.hctr_norm_open(arr, len(files), self.inp.label_dim, self.inp.dense_dim, bs, bh,
                                         len(blocks), so, self.key_np.itemsize,
                                         8 if self.key_dtype == torch.int64 else 4,
                                         int(self.check == Check_t.Sum), self.b * self.world, self.b,
                                         self.rank, int(self.repeat))
        vp = C.c_void_p
        ptrs = [(vp * self.DEPTH)(*[s[i].data_ptr() for s in self.slots]) for i in range(4)]
        self.lib.hctr_norm_start(self.h, self.DEPTH, *ptrs)
        self.started = True

    def read_a_batch(self):
        if self.h is None:
            self.start()
        valid = C.c_int(0)
        if getattr(self, "_last", None) is not None:     # its slot is recycled by the call below
            self._last.wait_copied()
            self._last = None
        idx = self.lib.hctr_norm_next(self.h, C.byref(valid))
        if valid.value == -2:
            msg = self.lib.hctr_norm_error(self.h).decode()
            raise (DataCheckError if "checksum" in msg else RuntimeError)(msg)
        if valid.value < 0:
            return None
        lab, den, keys, nnz = self.slots[idx]
        self.current_batchsize = valid.value
        mine = max(0, min(self.b, valid.value - self.rank * self.b))
        self._last = HostBatch(lab, den, keys[:self.nk], nnz[:self.nn], mine)
        return self._last

    def stop(self):
        if getattr(self, "h", None) is not None:
            self.lib.hctr_norm_close(self.h)
            self.h = None
            self.started = False

    def __del__(self):
        try:
            self.stop()
        except Exception:
            pass
