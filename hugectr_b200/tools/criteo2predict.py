"""Inference input files (the reference's tools/criteo_predict/criteo2predict.py format) and their loader.

A prediction file is four delimiter-separated lines: ``batch`` labels, ``batch x dense_dim`` dense values,
the embedding keys sample-major / slot-major, and the CSR row pointers (``batch x slots + 1``).  ``convert`` writes one
from a preprocessed text / CSV data set or from arrays; ``load`` reads it back into tensors that
``hugectr.inference.InferenceSession.predict`` takes.

  python -m hugectr_b200.tools.criteo2predict --src_csv_path test.txt --src_config_path dcn_data.json \\
         --dst_path dcn_input.txt --batch_size 128
"""
from __future__ import annotations

import json
from typing import Sequence

import numpy as np


def parse_config(path: str):
    j = json.load(open(path))
    dense, cat, slots = int(j["dense"]), int(j["categorical"]), [int(x) for x in j["slot_size"]]
    if cat != sum(slots):
        raise ValueError(f"categorical ({cat}) != sum(slot_size) ({sum(slots)})")
    return dense, cat, slots


def write(dst: str, label, dense, keys, slot_size: Sequence[int], sep: str = " "):
    """label [b], dense [b, D], keys [b, sum(slot_size)] (int) -> the four-line file"""
    label, dense, keys = np.asarray(label), np.asarray(dense), np.asarray(keys)
    b, ns = label.shape[0], len(slot_size)
    ptr = np.zeros(b * ns + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(np.tile(np.asarray(slot_size, dtype=np.int64), b))
    fmt = lambda a, f: sep.join(f % x for x in a.reshape(-1))
    with open(dst, "w") as f:
        f.write(fmt(label, "%g") + "\n")
        f.write(fmt(dense, "%.8g") + "\n")
        f.write(fmt(keys.astype(np.int64), "%d") + "\n")
        f.write(fmt(ptr, "%d") + "\n")


def convert(src_csv: str, src_config: str, dst: str, batch_size: int = 128, segmentation: str = " ",
            src_sep: str = " "):
    dense_dim, cat_dim, slots = parse_config(src_config)
    rows = np.loadtxt(src_csv, delimiter=None if src_sep == " " else src_sep, max_rows=batch_size, ndmin=2)
    if rows.shape[1] != 1 + dense_dim + cat_dim:
        raise ValueError(f"{src_csv}: {rows.shape[1]} columns, expected {1 + dense_dim + cat_dim}")
    write(dst, rows[:, 0], rows[:, 1:1 + dense_dim], rows[:, 1 + dense_dim:], slots, segmentation)
    return rows.shape[0]


def load(path: str, dense_dim: int, sep: str = " "):
    """-> (label [b] float32, dense [b, D] float32, keys [nnz] int64, row_ptrs [b * slots + 1] int64) tensors"""
    import torch
    with open(path) as f:
        lines = [l.strip() for l in f.read().splitlines() if l.strip()]
    if len(lines) != 4:
        raise ValueError(f"{path}: expected 4 lines, found {len(lines)}")
    tok = lambda l: l.split(sep) if sep != " " else l.split()
    label = np.asarray(tok(lines[0]), dtype=np.float32)
    dense = np.asarray(tok(lines[1]), dtype=np.float32).reshape(label.shape[0], dense_dim) if dense_dim else \
        np.zeros((label.shape[0], 0), np.float32)
    keys = np.asarray(tok(lines[2]), dtype=np.int64)
    ptr = np.asarray(tok(lines[3]), dtype=np.int64)
    if ptr[-1] != keys.shape[0]:
        raise ValueError(f"{path}: row pointers end at {ptr[-1]}, {keys.shape[0]} keys present")
    return tuple(torch.from_numpy(x) for x in (label, dense, keys, ptr))


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="Convert preprocessed Criteo data to the inference input format")
    ap.add_argument("--src_csv_path", required=True)
    ap.add_argument("--src_config_path", required=True)
    ap.add_argument("--dst_path", required=True)
    ap.add_argument("--batch_size", type=int, default=128)
    ap.add_argument("--segmentation", default=" ")
    a = ap.parse_args(argv)
    n = convert(a.src_csv_path, a.src_config_path, a.dst_path, a.batch_size, a.segmentation)
    print(f"{n} samples -> {a.dst_path}")
    return n


if __name__ == "__main__":
    main()
