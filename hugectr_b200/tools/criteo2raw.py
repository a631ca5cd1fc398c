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


class CriteoPreprocessor:
    """Native (multi-threaded C++) version of ``build_vocab`` + ``convert``:
    ``fit`` counts categorical values over any number of TSV files, ``finalize`` assigns ids by
    (count desc, value asc) with 0 = rare / unseen, ``transform`` writes the Raw binary
    (csrc/host/criteo_preprocess.cpp; reference tools/raw_script/criteo2raw.cpp, dlrm_script/dlrm_raw.cu)."""

    def __init__(self, num_dense: int = 13, num_cat: int = 26, num_threads: int = 8):
        import ctypes as C
        from .. import _native
        self.C = C
        self.lib = L = _native.host_lib()
        vp, ll = C.c_void_p, C.c_longlong
        L.hctr_criteo_open.restype = vp
        L.hctr_criteo_open.argtypes = [C.c_int, C.c_int]
        L.hctr_criteo_close.argtypes = [vp]
        L.hctr_criteo_fit.restype = ll
        L.hctr_criteo_fit.argtypes = [vp, C.c_char_p, C.c_int]
        L.hctr_criteo_finalize.argtypes = [vp, ll, ll]
        L.hctr_criteo_vocab_size.restype = ll
        L.hctr_criteo_vocab_size.argtypes = [vp, C.c_int]
        L.hctr_criteo_vocab_dump.argtypes = [vp, C.c_int, vp]
        L.hctr_criteo_vocab_load.argtypes = [vp, C.c_int, vp, ll]
        L.hctr_criteo_transform.restype = ll
        L.hctr_criteo_transform.argtypes = [vp, C.c_char_p, C.c_char_p, ll, ll, C.c_int]
        self.num_dense, self.num_cat, self.num_threads = num_dense, num_cat, num_threads
        self.h = L.hctr_criteo_open(num_dense, num_cat)
        self.num_lines = 0

    def fit(self, *tsv_paths):
        for p in tsv_paths:
            n = self.lib.hctr_criteo_fit(self.h, str(p).encode(), self.num_threads)
            if n < 0:
                raise FileNotFoundError(p)
            self.num_lines += n
        return self

    def finalize(self, min_freq: int = 1, max_size: int = 0):
        self.lib.hctr_criteo_finalize(self.h, int(min_freq), int(max_size))
        return self.vocab_sizes

    @property
    def vocab_sizes(self):
        return [int(self.lib.hctr_criteo_vocab_size(self.h, j)) for j in range(self.num_cat)]

    def vocabulary(self, col: int) -> np.ndarray:
        """values in id order: ``vocabulary(j)[i]`` has id ``i + 1``"""
        out = np.zeros(self.vocab_sizes[col] - 1, dtype=np.uint64)
        if out.size:
            self.lib.hctr_criteo_vocab_dump(self.h, col, out.ctypes.data)
        return out

    def load_vocabulary(self, col: int, values):
        v = np.ascontiguousarray(values, dtype=np.uint64)
        self.lib.hctr_criteo_vocab_load(self.h, col, v.ctypes.data, v.size)

    def transform(self, tsv_path: str, out_path: str, max_ind_range: int = 0, append_at: int = 0) -> int:
        """writes records starting at record index ``append_at``; returns the number written"""
        n = self.lib.hctr_criteo_transform(self.h, str(tsv_path).encode(), str(out_path).encode(),
                                           int(append_at), int(max_ind_range), self.num_threads)
        if n < 0:
            raise OSError(f"cannot convert {tsv_path} -> {out_path}")
        return int(n)

    def close(self):
        if self.h:
            self.lib.hctr_criteo_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="Criteo TSV -> Raw binary (frequency-thresholded ids)")
    ap.add_argument("--train", nargs="+", required=True, help="TSV files the vocabulary is fitted on")
    ap.add_argument("--convert", nargs="*", default=None, help="TSV files to convert (default: --train)")
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--min-freq", type=int, default=1)
    ap.add_argument("--max-ind-range", type=int, default=0)
    ap.add_argument("--threads", type=int, default=8)
    a = ap.parse_args(argv)
    import os
    os.makedirs(a.out_dir, exist_ok=True)
    pre = CriteoPreprocessor(num_threads=a.threads).fit(*a.train)
    sizes = pre.finalize(a.min_freq)
    for f in (a.convert or a.train):
        out = os.path.join(a.out_dir, os.path.basename(f) + ".bin")
        print(out, pre.transform(f, out, a.max_ind_range), "records")
    print("slot_size_array =", sizes if a.max_ind_range <= 0 else [min(s, a.max_ind_range) for s in sizes])
    return sizes


if __name__ == "__main__":
    main()
