"""Norm-format reader with optional per-record checksum verification (Check_t.Sum).

File: DataSetHeader {error_check, number_of_records, label_dim, dense_dim, slot_num, reserved[3]}
(8 x int64, HugeCTR/include/common.hpp:184-191) followed by records
[label f32 x L][dense f32 x D] then per slot {int nnz, keys[nnz]}; with check_sum every record is
wrapped as {int nbytes, payload, char sum} (HugeCTR/include/data_generator.hpp:137-188,
include/data_readers/check_sum.hpp).  The reader side of component C46.
"""
from __future__ import annotations

import struct

import numpy as np
import torch

from ..enums import Check_t
from .batch import HostBatch
from .parquet_reader import read_file_list
from .readers import IDataReader


class DataCheckError(RuntimeError):
    pass


class NormReader(IDataReader):
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
