"""MLPerf DLRM-DCNv2 data set conversion: TorchRec NumPy days -> the raw binary the RawAsync reader trains
from (the role of samples/dlrm/preprocessing/convert_to_raw.py of the reference).

Input per day d: ``day_d_labels.npy`` [n], ``day_d_dense.npy`` [n, 13], ``day_d_sparse_multi_hot.npz``
with arrays "0".."25" of shape [n, hotness_i].  Output ``{train,val,test}_data.bin``: one fixed record per
sample = label bytes | dense bytes | feature 0 keys | ... | feature 25 keys; train = days 0..22, val / test =
the last day split at sample 89,137,319.

Unlike the reference script (which loads a whole day, ~200 GB of RAM, and writes row by row from Python)
the inputs are memory-mapped -- members of an uncompressed ``.npz`` are mapped in place through their zip
offsets -- and records are assembled a chunk at a time with vectorised byte-plane copies, so memory stays
at one chunk (default 1 M samples ~ 0.9 GB) and the speed is that of the storage.

    python -m hugectr_b200.tools.convert_to_raw --input_dir_labels_and_dense in/ \\
        --input_dir_sparse_multihot in/ --output_dir out/ [--stages train val test]
"""
from __future__ import annotations

import argparse
import os
import struct
import time
import zipfile

import numpy as np

NUM_DAYS = 24
NUM_SPARSE = 26
LAST_DAY_TEST_VAL_SPLIT_POINT = 89_137_319
STAGES = ("train", "val", "test")


def npz_member(path: str, name: str) -> np.ndarray:
    """array ``name`` of an ``.npz``: memory-mapped in place when the member is stored uncompressed
    (``np.savez``), loaded when it is deflated (``np.savez_compressed``)"""
    member = name if name.endswith(".npy") else name + ".npy"
    with zipfile.ZipFile(path) as z:
        info = z.getinfo(member)
        if info.compress_type != zipfile.ZIP_STORED:
            with z.open(member) as f:
                return np.lib.format.read_array(f)
        with open(path, "rb") as f:
            f.seek(info.header_offset)
            hdr = f.read(30)
            n_name, n_extra = struct.unpack("<HH", hdr[26:30])
            f.seek(info.header_offset + 30 + n_name + n_extra)
            major, _ = np.lib.format.read_magic(f)
            shape, fortran, dtype = (np.lib.format.read_array_header_1_0(f) if major == 1
                                     else np.lib.format.read_array_header_2_0(f))
            assert not fortran, "C-like index order expected"
            off = f.tell()
    return np.memmap(path, dtype=dtype, mode="r", offset=off, shape=shape)


def write_records(out, label, dense, sparse, start: int, stop: int, chunk: int) -> int:
    """append samples [start, stop) as raw records; returns the number written"""
    lab2 = label.reshape(label.shape[0], -1)
    if lab2.dtype.kind == "f":
        # bytes are kept as they are (as the reference's converter does); the MLPerf numpy dumps carry int32
        # labels, which is what the raw reader with is_dense_float expects (split_batch.cu:43-88)
        import warnings
        warnings.warn("float labels: read the raw file with DataReaderParams(float_label_dense=True), or store "
                      "the labels as int32", stacklevel=2)
    parts = [lab2, dense.reshape(dense.shape[0], -1)] + [s.reshape(s.shape[0], -1) for s in sparse]
    widths = [p.shape[1] * p.dtype.itemsize for p in parts]
    rec_bytes = sum(widths)
    done = 0
    for lo in range(start, stop, chunk):
        hi = min(stop, lo + chunk)
        m = hi - lo
        rec = np.empty((m, rec_bytes), dtype=np.uint8)
        off = 0
        for p, w in zip(parts, widths):
            rec[:, off:off + w] = np.ascontiguousarray(p[lo:hi]).view(np.uint8).reshape(m, w)
            off += w
        rec.tofile(out)
        done += m
    return done


def convert(input_dir_labels_and_dense: str, input_dir_sparse_multihot: str, output_dir: str,
            stages=STAGES, chunk_size: int = 1 << 20, num_days: int = NUM_DAYS,
            split_point: int = LAST_DAY_TEST_VAL_SPLIT_POINT, log=print) -> dict:
    os.makedirs(output_dir, exist_ok=True)
    counts = {}
    for stage in stages:
        days = list(range(num_days - 1)) if stage == "train" else [num_days - 1]
        path = os.path.join(output_dir, f"{stage}_data.bin")
        t0, total = time.perf_counter(), 0
        with open(path, "wb") as out:
            for d in days:
                label = np.load(os.path.join(input_dir_labels_and_dense, f"day_{d}_labels.npy"), mmap_mode="r")
                dense = np.load(os.path.join(input_dir_labels_and_dense, f"day_{d}_dense.npy"), mmap_mode="r")
                npz = os.path.join(input_dir_sparse_multihot, f"day_{d}_sparse_multi_hot.npz")
                sparse = [npz_member(npz, str(i)) for i in range(NUM_SPARSE)]
                n = label.shape[0]
                assert dense.shape[0] == n and all(s.shape[0] == n for s in sparse), f"day {d}: row counts differ"
                lo, hi = 0, n
                if stage == "val":
                    hi = min(n, split_point)
                elif stage == "test":
                    lo = min(n, split_point)
                total += write_records(out, label, dense, sparse, lo, hi, chunk_size)
                log(f"[{stage}] day {d}: {hi - lo:,} samples")
        dt = time.perf_counter() - t0
        counts[stage] = total
        log(f"[{stage}] {total:,} samples -> {path} ({total / max(dt, 1e-9):,.0f} samples/s)")
    return counts


def main(argv=None):
    ap = argparse.ArgumentParser(description="NumPy to Raw format conversion (MLPerf DLRM-DCNv2 data set)")
    ap.add_argument("--input_dir_labels_and_dense", required=True)
    ap.add_argument("--input_dir_sparse_multihot", required=True)
    ap.add_argument("--output_dir", required=True)
    ap.add_argument("--stages", nargs="+", choices=STAGES, default=list(STAGES))
    ap.add_argument("--chunk_size", type=int, default=1 << 20)
    ap.add_argument("--num_days", type=int, default=NUM_DAYS)
    ap.add_argument("--split_point", type=int, default=LAST_DAY_TEST_VAL_SPLIT_POINT)
    a = ap.parse_args(argv)
    return convert(a.input_dir_labels_and_dense, a.input_dir_sparse_multihot, a.output_dir, a.stages,
                   a.chunk_size, a.num_days, a.split_point)


if __name__ == "__main__":
    main()
