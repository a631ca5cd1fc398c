"""criteo2raw / dlrm_raw equivalents: convert Criteo-style TSV (label, 13 ints, 26 hex categoricals)
to the RawAsync binary layout, with frequency-thresholded categorification
(tools/criteo_script, tools/dlrm_script/dlrm_raw.cu).  CPU/numpy implementation."""
from __future__ import annotations

import numpy as np


def build_vocab(columns, min_freq: int = 1, max_size: int = 0):
    """per-column {value -> dense id}; id 0 is reserved for out-of-vocabulary / rare values"""
    vocabs = []
    for col in columns:
        vals, cnt = np.unique(col, return_counts=True)
        keep = vals[cnt >= min_freq]
        order = np.argsort(-cnt[cnt >= min_freq], kind="stable")
        keep = keep[order]
        if max_size > 0:
            keep = keep[:max_size - 1]
        vocabs.append({v: i + 1 for i, v in enumerate(keep.tolist())})
    return vocabs


def convert(tsv_path: str, out_path: str, vocabs=None, min_freq: int = 1, max_ind_range: int = 0,
            num_dense: int = 13, num_cat: int = 26):
    labels, dense, cats = [], [], [[] for _ in range(num_cat)]
    with open(tsv_path) as f:
        for line in f:
            p = line.rstrip("\n").split("\t")
            labels.append(int(p[0] or 0))
            dense.append([max(0, int(x)) if x else 0 for x in p[1:1 + num_dense]])
            for j in range(num_cat):
                v = p[1 + num_dense + j]
                cats[j].append(int(v, 16) if v else 0)
    cats = [np.asarray(c, dtype=np.int64) for c in cats]
    if vocabs is None:
        vocabs = build_vocab(cats, min_freq, max_ind_range)
    n = len(labels)
    ids = np.zeros((n, num_cat), dtype="<u4")
    for j in range(num_cat):
        m = vocabs[j]
        ids[:, j] = [m.get(int(v), 0) for v in cats[j]]
        if max_ind_range > 0:
            ids[:, j] %= max_ind_range
    rec = np.concatenate([np.asarray(labels, dtype="<i4").reshape(-1, 1).view("<u4"),
                          np.asarray(dense, dtype="<u4"), ids], 1)
    rec.tofile(out_path)
    return vocabs, [len(v) + 1 for v in vocabs]
