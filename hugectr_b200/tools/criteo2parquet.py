"""Criteo TSV -> categorified Parquet data set for the Parquet reader.

The role of the reference's NVTabular script (tools/criteo_script/preprocess_nvt.py: fill missing values, clip
negative dense values, optional log(x + 1), frequency-thresholded Categorify, ``_metadata.json`` + file list per split)
on this stack's native preprocessing: the vocabulary fit and the per-line transform are the multi-threaded C++ of
``tools.criteo2raw.CriteoPreprocessor`` (csrc/host/criteo_preprocess.cpp); the fixed-width records it writes are then
re-columnised into Parquet row groups with pyarrow.

  python -m hugectr_b200.tools.criteo2parquet --train day_0 day_1 --val day_2 --out-dir criteo_parquet \\
         --freq-limit 6 --normalize-dense

Output: ``<out>/train/*.parquet``, ``<out>/val/*.parquet``, each with ``_metadata.json`` (labels / conts / cats by
column index, rows per file) and ``_file_list.txt``; prints the ``slot_size_array`` to pass to ``DataReaderParams``.
"""
from __future__ import annotations

import json
import os
import tempfile
from typing import List, Sequence

import numpy as np

from .criteo2raw import CriteoPreprocessor


def _write_split(pre: CriteoPreprocessor, tsvs: Sequence[str], out_dir: str, rows_per_file: int,
                 normalize_dense: bool, max_ind_range: int, row_group: int) -> int:
    import pyarrow as pa
    import pyarrow.parquet as pq
    os.makedirs(out_dir, exist_ok=True)
    nd, nc = pre.num_dense, pre.num_cat
    width = 1 + nd + nc
    names = ["label"] + [f"I{i + 1}" for i in range(nd)] + [f"C{i + 1}" for i in range(nc)]
    stats, paths, total = [], [], 0
    with tempfile.TemporaryDirectory(dir=out_dir) as tmp:
        for ti, tsv in enumerate(tsvs):
            raw = os.path.join(tmp, "part.bin")
            n = pre.transform(tsv, raw, max_ind_range)
            rec = np.memmap(raw, dtype="<u4", mode="r", shape=(n, width)) if n else np.zeros((0, width), "<u4")
            for fi, lo in enumerate(range(0, n, rows_per_file)):
                hi = min(n, lo + rows_per_file)
                blk = np.asarray(rec[lo:hi])
                dense = blk[:, 1:1 + nd].view("<i4").astype(np.float32)
                if normalize_dense:
                    dense = np.log1p(np.maximum(dense, 0.0)).astype(np.float32)
                cols = [pa.array(blk[:, 0].view("<i4").astype(np.float32))]
                cols += [pa.array(np.ascontiguousarray(dense[:, i])) for i in range(nd)]
                cols += [pa.array(blk[:, 1 + nd + j].astype(np.int64)) for j in range(nc)]
                name = f"{os.path.basename(tsv)}.{fi}.parquet" if n > rows_per_file else f"{os.path.basename(tsv)}.parquet"
                path = os.path.join(out_dir, name)
                pq.write_table(pa.Table.from_arrays(cols, names=names), path, row_group_size=row_group)
                stats.append({"file_name": name, "num_rows": int(hi - lo)})
                paths.append(path)
            total += n
            del rec
    meta = {"file_stats": stats,
            "labels": [{"col_name": "label", "index": 0}],
            "conts": [{"col_name": names[1 + i], "index": 1 + i} for i in range(nd)],
            "cats": [{"col_name": names[1 + nd + j], "index": 1 + nd + j} for j in range(nc)]}
    with open(os.path.join(out_dir, "_metadata.json"), "w") as f:
        json.dump(meta, f)
    with open(os.path.join(out_dir, "_file_list.txt"), "w") as f:
        f.write(f"{len(paths)}\n" + "\n".join(paths) + "\n")
    return total


def convert(train: Sequence[str], val: Sequence[str], out_dir: str, freq_limit: int = 1, max_ind_range: int = 0,
            normalize_dense: bool = False, rows_per_file: int = 1 << 22, row_group: int = 1 << 16,
            threads: int = 8, num_dense: int = 13, num_cat: int = 26, log=print) -> List[int]:
    """fit on ``train``, write ``<out_dir>/train`` and ``<out_dir>/val``; returns the slot_size_array"""
    pre = CriteoPreprocessor(num_dense, num_cat, threads).fit(*train)
    sizes = pre.finalize(freq_limit)
    if max_ind_range > 0:
        sizes = [min(s, max_ind_range) for s in sizes]
    for split, files in (("train", train), ("val", val)):
        if files:
            n = _write_split(pre, files, os.path.join(out_dir, split), rows_per_file, normalize_dense,
                             max_ind_range, row_group)
            log(f"{split}: {n} rows -> {os.path.join(out_dir, split)}")
    pre.close()
    log("slot_size_array =", sizes)
    return sizes


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="Criteo TSV -> categorified Parquet (+ _metadata.json, file lists)")
    ap.add_argument("--train", nargs="+", required=True)
    ap.add_argument("--val", nargs="*", default=[])
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--freq-limit", type=int, default=6, help="values seen fewer times map to id 0 (NVT freq_threshold)")
    ap.add_argument("--max-ind-range", type=int, default=0, help="ids are taken modulo this range when > 0")
    ap.add_argument("--normalize-dense", action="store_true", help="log(x + 1) of the clipped dense features")
    ap.add_argument("--rows-per-file", type=int, default=1 << 22)
    ap.add_argument("--threads", type=int, default=8)
    a = ap.parse_args(argv)
    return convert(a.train, a.val, a.out_dir, a.freq_limit, a.max_ind_range, a.normalize_dense, a.rows_per_file,
                   threads=a.threads)


if __name__ == "__main__":
    main()
