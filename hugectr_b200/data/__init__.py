"""Data pipeline: DataGenerator, readers (Parquet / RawAsync / Norm / synthetic), checkers."""
from __future__ import annotations

import os

import torch

from ..enums import DataReaderType_t
from .batch import HostBatch, power_law_keys
from .readers import IDataReader, SparseLayout, SyntheticReader


def _slot_vocab(model):
    """Per-slot key ranges: slot_size_array if given, else EBC table sizes, else a default."""
    rp = model.reader_params
    n = model.layout.total_slots
    if rp.slot_size_array and len(rp.slot_size_array) >= n:
        return [int(v) for v in rp.slot_size_array[:n]]
    vocab = {}
    for cfg in model.ebc_configs:
        for lk in cfg.lookups:
            for t, b in zip(lk["tables"], lk["bottoms"]):
                vocab[b] = t.max_vocabulary_size
    out = []
    for (name, S, H, fixed) in model.layout.blocks:
        for _ in range(S):
            out.append(int(vocab.get(name, 10000)))
    return out


def create_reader(model, is_train: bool) -> IDataReader:
    rp = model.reader_params
    inp = model.input
    b = model.b_train if is_train else model.b_eval
    rank, world = model.comm.rank, model.world
    src = rp.source if is_train else ([rp.eval_source] if rp.eval_source else [])
    synthetic = (not src) or all(str(s).startswith("synthetic") for s in src) or \
        os.environ.get("HCTR_FORCE_SYNTHETIC", "0") == "1"
    if synthetic:
        alpha = 1.1
        for s in src:
            if ":" in str(s):
                alpha = float(str(s).split(":")[1])
        return SyntheticReader(b, rank, world, inp.label_dim, inp.dense_dim, model.layout,
                               _slot_vocab(model), alpha=alpha,
                               seed=model.solver.seed + (0 if is_train else 1),
                               pool=int(os.environ.get("HCTR_SYNTH_POOL", "8")),
                               key_dtype=model.key_dtype)
    t = rp.data_reader_type
    if t == DataReaderType_t.Parquet:
        from .parquet_reader import ParquetReader
        return ParquetReader(model, is_train)
    if t == DataReaderType_t.RawAsync or t == DataReaderType_t.Raw:
        from .raw_reader import RawAsyncReader
        return RawAsyncReader(model, is_train)
    if t == DataReaderType_t.Norm:
        from .norm_reader import NormReader
        return NormReader(model, is_train)
    raise ValueError(f"unsupported reader type {t}")


class DataSource:
    """hugectr.data.DataSource of older releases: stages remote files next to the job
    (``move_to_local``) through the FileSystem layer (notebooks/training_with_remote_filesystem)."""

    def __init__(self, data_source_params, remote_paths=None, local_paths=None):
        self.params = data_source_params
        self.remote_paths = list(remote_paths or getattr(data_source_params, "filesystem_paths", []) or [])
        self.local_paths = list(local_paths or getattr(data_source_params, "local_paths", []) or [])

    def move_to_local(self):
        from ..io import FileSystemBuilder
        for src, dst in zip(self.remote_paths, self.local_paths):
            fs = FileSystemBuilder.build_by_path(src, self.params)
            os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
            fs.fetch(src, dst)
        return self.local_paths


def __getattr__(name):
    if name == "DataSourceParams":
        from ..solver import DataSourceParams
        return DataSourceParams
    if name in ("DataGenerator", "DataGeneratorParams"):
        from . import generator
        return getattr(generator, name)
    raise AttributeError(name)
