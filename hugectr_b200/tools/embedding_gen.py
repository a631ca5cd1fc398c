"""Pre-generate embedding weights on disk (the role of tools/model_generation/embedding_gen.py of the
reference, which writes the retired single-file localized format): writes an embedding-collection
checkpoint folder ``<out>/embedding_collection_0/{meta_data,key<t>,weight<t>}`` with keys 0..n-1 per table
and U(-1/sqrt(n), 1/sqrt(n)) rows, streamed in chunks, loadable with ``model.embedding_load(<out>)``.

    python -m hugectr_b200.tools.embedding_gen --embedding-size 39884406-39043-17289 --dim 128 --output ckpt/
"""
from __future__ import annotations

import argparse
import math
import os

import numpy as np

from ..io.checkpoint import _file_head


def generate(table_sizes, dim, out_dir: str, i64_keys: bool = True, seed: int = 0, chunk_rows: int = 1 << 20):
    dims = [int(dim)] * len(table_sizes) if isinstance(dim, int) else [int(d) for d in dim]
    folder = os.path.join(out_dir, "embedding_collection_0")
    os.makedirs(folder, exist_ok=True)
    kd = "<i8" if i64_keys else "<u4"
    rng = np.random.default_rng(seed)
    for t, (n, ev) in enumerate(zip(table_sizes, dims)):
        n = int(n)
        bound = math.sqrt(1.0 / max(n, 1))
        with open(os.path.join(folder, f"key{t}"), "wb") as fk, open(os.path.join(folder, f"weight{t}"), "wb") as fw:
            fk.write(_file_head(1, t))
            fw.write(_file_head(2, t))
            for lo in range(0, n, chunk_rows):
                hi = min(n, lo + chunk_rows)
                np.arange(lo, hi, dtype=kd).tofile(fk)
                rng.uniform(-bound, bound, (hi - lo, ev)).astype("<f4").tofile(fw)
    head = np.zeros(5, dtype="<i4")
    head[0], head[1] = len(table_sizes), 1 if i64_keys else 0
    meta = head.tobytes() + np.arange(len(table_sizes), dtype="<i4").tobytes() + \
        np.asarray(table_sizes, dtype="<u8").tobytes() + np.asarray(dims, dtype="<i4").tobytes()
    with open(os.path.join(folder, "meta_data"), "wb") as f:
        f.write(meta)
    return folder


def main(argv=None):
    ap = argparse.ArgumentParser(description="Pre-generate embedding weights")
    ap.add_argument("--embedding-size", type=str, required=True, help="table sizes separated by '-'")
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--output", type=str, required=True)
    ap.add_argument("--u32-keys", action="store_true")
    a = ap.parse_args(argv)
    sizes = [int(x) for x in a.embedding_size.split("-")]
    print("Embedding size:", sizes, "-> ", generate(sizes, a.dim, a.output, not a.u32_keys))


if __name__ == "__main__":
    main()
