"""hugectr.tools.DataGenerator: synthetic datasets in Norm / Raw / Parquet format.

Parity: HugeCTR/include/pybind/data_generator_wrapper.hpp:29-70, HugeCTR/src/data_generator.cpp:31-295,
HugeCTR/include/data_generator.hpp (Norm :189-330 with DataSetHeader + optional per-record checksum,
Raw :978-1070 one binary file of fixed records [label][dense][keys], Parquet :500-660 with file list
"<num_files>\\n<path>..." and _metadata.json {file_stats, labels, conts, cats}).  Key distribution:
uniform or power law (alpha Long .9 / Medium 1.1 / Short 1.3 / Specific), inverse-CDF sampler :109-131.

Norm and Raw files are written by the native generator (csrc/host/data_generator.cpp): counter-based
per-record random streams (the bytes do not depend on ``num_threads``), one thread per Norm file /
per 8 Ki-record chunk of a Raw file.  ``HCTR_DATAGEN=python`` selects the numpy writers below.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import struct
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from .. import _native
from ..enums import Check_t, DataReaderType_t, Distribution_t, PowerLaw_t
from ..utils import logger
from .batch import power_law_keys

_ALPHA = {PowerLaw_t.Long: 0.9, PowerLaw_t.Medium: 1.1, PowerLaw_t.Short: 1.3}


@dataclass
class DataGeneratorParams:
    format: DataReaderType_t
    label_dim: int
    dense_dim: int
    num_slot: int
    i64_input_key: bool
    source: str
    eval_source: str
    slot_size_array: List[int]
    nnz_array: List[int] = field(default_factory=list)
    check_type: Check_t = Check_t.Sum
    dist_type: Distribution_t = Distribution_t.PowerLaw
    power_law_type: PowerLaw_t = PowerLaw_t.Specific
    alpha: float = 1.2
    num_files: int = 128
    eval_num_files: int = 32
    num_samples_per_file: int = 40960
    num_samples: int = 5242880
    eval_num_samples: int = 1310720
    float_label_dense: bool = False
    num_threads: int = 1

    def __post_init__(self):
        if len(self.slot_size_array) != self.num_slot:
            raise ValueError("slot_size_array.size() should be equal to num_slot")
        if self.nnz_array and len(self.nnz_array) != self.num_slot:
            raise ValueError("nnz_array.size() should be equal to num_slot")


class DataGenerator:
    def __init__(self, data_generator_params: DataGeneratorParams):
        self.p = data_generator_params
        self.gen = torch.Generator()
        self.gen.manual_seed(20260921)

    @property
    def alpha(self):
        p = self.p
        return p.alpha if p.power_law_type == PowerLaw_t.Specific else _ALPHA[p.power_law_type]

    def _keys(self, n, vocab, gen=None):
        gen = gen or self.gen
        if self.p.dist_type == Distribution_t.PowerLaw:
            return power_law_keys(n, vocab, self.alpha, gen).numpy()
        return torch.randint(0, max(1, vocab), (n,), generator=gen).numpy()

    def generate(self):
        p = self.p
        logger.info(f"Generate {p.format.name} dataset: {p.source} / {p.eval_source}")
        if p.format == DataReaderType_t.Parquet:
            self._parquet(p.source, p.num_files, "train")
            self._parquet(p.eval_source, p.eval_num_files, "val")
        elif p.format == DataReaderType_t.Norm:
            self._norm(p.source, p.num_files, "train")
            self._norm(p.eval_source, p.eval_num_files, "val")
        elif p.format in (DataReaderType_t.Raw, DataReaderType_t.RawAsync):
            self._raw(p.source, p.num_samples)
            self._raw(p.eval_source, p.eval_num_samples)
        else:
            raise ValueError(f"unsupported format {p.format}")

    # ------------------------------------------------------------------ Parquet
    def _parquet(self, file_list: str, num_files: int, sub: str):
        import pyarrow as pa
        import pyarrow.parquet as pq
        p = self.p
        root = os.path.dirname(file_list) or "."
        d = os.path.join(root, sub)
        os.makedirs(d, exist_ok=True)
        n = p.num_samples_per_file
        names = []
        for i in range(p.label_dim):
            names.append(f"label{i}" if p.label_dim > 1 else "label")
        names += [f"C{i + 1}" for i in range(p.dense_dim)] + [f"S{s + 1}" for s in range(p.num_slot)]

        def one_file(f):
            # every file has its own random stream (seeded by split and file index): files are written in
            # parallel and the bytes do not depend on the number of writer threads
            gen = torch.Generator()
            gen.manual_seed(20260921 + 1000003 * (f + 1) + (0 if sub == "train" else 7919))
            cols = []
            for i in range(p.label_dim):
                cols.append(pa.array(torch.rand(n, generator=gen).round().numpy().astype("float32")))
            for i in range(p.dense_dim):
                cols.append(pa.array(torch.rand(n, generator=gen).numpy().astype("float32")))
            for s in range(p.num_slot):
                nnz = p.nnz_array[s] if p.nnz_array else 1
                vocab = p.slot_size_array[s]
                if nnz == 1:
                    cols.append(pa.array(self._keys(n, vocab, gen).astype("int64")))
                else:
                    cnt = torch.randint(1, nnz + 1, (n,), generator=gen).numpy()
                    flat = self._keys(int(cnt.sum()), vocab, gen).astype("int64")
                    offs = np.concatenate([[0], np.cumsum(cnt)]).astype("int32")
                    cols.append(pa.ListArray.from_arrays(pa.array(offs), pa.array(flat)))
            name = f"gen_{f}.parquet"
            path = os.path.join(d, name)
            pq.write_table(pa.Table.from_arrays(cols, names=names), path, row_group_size=min(n, 65536))
            return {"file_name": name, "num_rows": n}, path
        from concurrent.futures import ThreadPoolExecutor
        workers = max(1, min(int(getattr(p, "num_threads", 0) or 8), num_files, os.cpu_count() or 1))
        with ThreadPoolExecutor(max_workers=workers) as ex:
            done = list(ex.map(one_file, range(num_files)))
        stats, paths = [x[0] for x in done], [x[1] for x in done]
        nl, nd = p.label_dim, p.dense_dim
        meta = {"file_stats": stats,
                "labels": [{"col_name": names[i], "index": i} for i in range(nl)],
                "conts": [{"col_name": names[nl + i], "index": nl + i} for i in range(nd)],
                "cats": [{"col_name": names[nl + nd + i], "index": nl + nd + i} for i in range(p.num_slot)]}
        with open(os.path.join(d, "_metadata.json"), "w") as fo:
            json.dump(meta, fo)
        with open(file_list, "w") as fo:
            fo.write(f"{num_files}\n" + "\n".join(paths) + "\n")

    # ------------------------------------------------------------------ native writers
    SEED = 20260921

    def _native_args(self):
        p = self.p
        L = _native.host_lib()
        ll, ip = C.POINTER(C.c_longlong), C.POINTER(C.c_int)
        L.hctr_gen_norm_file.argtypes = [C.c_char_p, C.c_ulonglong, C.c_longlong, C.c_longlong, C.c_int,
                                         C.c_int, C.c_int, ll, ip, C.c_int, C.c_int, C.c_int, C.c_double]
        L.hctr_gen_raw_file.argtypes = [C.c_char_p, C.c_ulonglong, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                        ll, ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
        sizes = (C.c_longlong * p.num_slot)(*[int(v) for v in p.slot_size_array])
        nnz = (C.c_int * p.num_slot)(*[int(v) for v in (p.nnz_array or [1] * p.num_slot)])
        return L, sizes, nnz, 8 if p.i64_input_key else 4, int(p.dist_type == Distribution_t.PowerLaw)

    def _norm(self, file_list: str, num_files: int, sub: str):
        if os.environ.get("HCTR_DATAGEN", "native") == "python":
            return self._norm_py(file_list, num_files, sub)
        p = self.p
        d = os.path.join(os.path.dirname(file_list) or ".", sub)
        os.makedirs(d, exist_ok=True)
        L, sizes, nnz, kb, pl = self._native_args()
        paths = [os.path.join(d, f"gen_{f}.data") for f in range(num_files)]
        fid0 = 0 if sub == "train" else 1 << 32      # train / eval draw from different streams

        def one(f):
            return L.hctr_gen_norm_file(paths[f].encode(), self.SEED, fid0 + f, p.num_samples_per_file,
                                        p.label_dim, p.dense_dim, p.num_slot, sizes, nnz, kb,
                                        int(p.check_type == Check_t.Sum), pl, float(self.alpha))
        with ThreadPoolExecutor(max_workers=max(1, min(p.num_threads, num_files or 1))) as ex:
            rcs = list(ex.map(one, range(num_files)))
        if any(rcs):
            raise OSError(f"cannot write Norm files under {d}")
        with open(file_list, "w") as fo:
            fo.write(f"{num_files}\n" + "\n".join(paths) + "\n")

    def _raw(self, path: str, num_samples: int):
        """fixed records [label_dim x (f32|i32)][dense_dim x (f32|i32)][sum(nnz) x key]"""
        if os.environ.get("HCTR_DATAGEN", "native") == "python":
            return self._raw_py(path, num_samples)
        p = self.p
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        L, sizes, nnz, kb, pl = self._native_args()
        seed = self.SEED + (0 if path == p.source else 1)
        rc = L.hctr_gen_raw_file(path.encode(), seed, int(num_samples), p.label_dim, p.dense_dim,
                                 p.num_slot, sizes, nnz, kb, int(p.float_label_dense), pl,
                                 float(self.alpha), max(1, p.num_threads))
        if rc:
            raise OSError(f"cannot write {path}")

    # ------------------------------------------------------------------ Norm (numpy writer)
    def _norm_py(self, file_list: str, num_files: int, sub: str):
        p = self.p
        root = os.path.dirname(file_list) or "."
        d = os.path.join(root, sub)
        os.makedirs(d, exist_ok=True)
        kt = "<i8" if p.i64_input_key else "<u4"
        paths = []
        for f in range(num_files):
            path = os.path.join(d, f"gen_{f}.data")
            n = p.num_samples_per_file
            with open(path, "wb") as fo:
                chk = 1 if p.check_type == Check_t.Sum else 0
                fo.write(struct.pack("<8q", chk, n, p.label_dim, p.dense_dim, p.num_slot, 0, 0, 0))
                labels = torch.rand(n, p.label_dim, generator=self.gen).round().numpy().astype("<f4")
                dense = torch.rand(n, p.dense_dim, generator=self.gen).numpy().astype("<f4")
                slot_keys, slot_cnt = [], []
                for s in range(p.num_slot):
                    nnz = p.nnz_array[s] if p.nnz_array else 1
                    cnt = np.full(n, 1, dtype="<i4") if nnz == 1 else \
                        torch.randint(1, nnz + 1, (n,), generator=self.gen).numpy().astype("<i4")
                    slot_cnt.append(cnt)
                    slot_keys.append(self._keys(int(cnt.sum()), p.slot_size_array[s]).astype(kt))
                pos = [0] * p.num_slot
                for i in range(n):
                    rec = bytearray()
                    rec += labels[i].tobytes() + dense[i].tobytes()
                    for s in range(p.num_slot):
                        c = int(slot_cnt[s][i])
                        rec += struct.pack("<i", c) + slot_keys[s][pos[s]:pos[s] + c].tobytes()
                        pos[s] += c
                    if chk:
                        fo.write(struct.pack("<i", len(rec)) + rec +
                                 struct.pack("<b", np.frombuffer(bytes(rec), dtype=np.int8).sum(dtype=np.int8)))
                    else:
                        fo.write(rec)
            paths.append(path)
        with open(file_list, "w") as fo:
            fo.write(f"{num_files}\n" + "\n".join(paths) + "\n")

    # ------------------------------------------------------------------ Raw (numpy writer)
    def _raw_py(self, path: str, num_samples: int):
        p = self.p
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        nnz = p.nnz_array or [1] * p.num_slot
        kt = np.dtype("<i8") if p.i64_input_key else np.dtype("<u4")
        chunk = 1 << 16
        with open(path, "wb") as fo:
            done = 0
            while done < num_samples:
                n = min(chunk, num_samples - done)
                lab = torch.rand(n, p.label_dim, generator=self.gen).round().numpy()
                den = torch.rand(n, p.dense_dim, generator=self.gen).numpy()
                if p.float_label_dense:
                    ld = np.concatenate([lab.astype("<f4"), den.astype("<f4")], 1).view("<u4")
                else:
                    ld = np.concatenate([lab.astype("<i4"), (den * 100).astype("<i4")], 1).view("<u4")
                keys = np.concatenate([self._keys(n * nnz[s], p.slot_size_array[s]).reshape(n, nnz[s])
                                       for s in range(p.num_slot)], 1).astype(kt)
                if kt.itemsize == 4:
                    rec = np.concatenate([ld, keys.view("<u4")], 1)
                    fo.write(rec.tobytes())
                else:
                    for i in range(n):
                        fo.write(ld[i].tobytes() + keys[i].tobytes())
                done += n
