#!/usr/bin/env python
"""Reader throughput: generate a Criteo-shaped synthetic data set in each on-disk format (tools.DataGenerator)
and drain it through the framework's own readers -- the path `model.train()` pulls batches from (decode /
split into label, dense and feature-major keys inside pinned ring slots; on a GPU also the H2D copy and, for
RawAsync, the device split kernel).

  python benchmarks/data_readers/benchmark.py [--formats parquet,raw,norm] [--samples 2000000] [--batch 16384]

Prints one JSON line per format: samples/s, file MB/s, the reader's thread count.
(reference counterpart: the data-reader unit benchmarks under test/utest/data_reader/ and the figures quoted in
docs/source/performance.md -- no published numbers for these readers, so the line is a measurement, not a ratio)
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch  # noqa: E402

import hugectr_b200 as hugectr  # noqa: E402
from hugectr_b200.data.generator import DataGenerator, DataGeneratorParams  # noqa: E402

SLOTS = [1461, 558, 335378, 211710, 306, 20, 12136, 634, 4, 51298, 5302, 332600, 3179, 27, 12191,
         301211, 11, 4841, 2086, 4, 324273, 17, 16, 79734, 96, 58622]       # Criteo Kaggle cardinalities


def dir_bytes(paths):
    tot = 0
    for p in paths:
        if os.path.isdir(p):
            for r, _, fs in os.walk(p):
                tot += sum(os.path.getsize(os.path.join(r, f)) for f in fs)
        elif os.path.exists(p):
            tot += os.path.getsize(p)
    return tot


def build(fmt, root, n, batch, threads):
    T = hugectr.DataReaderType_t
    kind = {"parquet": T.Parquet, "raw": T.RawAsync, "norm": T.Norm}[fmt]
    d = os.path.join(root, fmt)
    os.makedirs(d, exist_ok=True)
    if fmt == "parquet":
        src, ev = os.path.join(d, "file_list.txt"), os.path.join(d, "file_list_test.txt")
    elif fmt == "raw":
        src, ev = os.path.join(d, "train.bin"), os.path.join(d, "val.bin")
    else:
        src, ev = os.path.join(d, "train.txt"), os.path.join(d, "val.txt")
    nfiles = 8
    gp = DataGeneratorParams(kind, 1, 13, 26, fmt != "raw", src, ev, SLOTS, nnz_array=[1] * 26,
                             num_files=nfiles, eval_num_files=1, num_samples_per_file=n // nfiles,
                             num_samples=n, eval_num_samples=batch, num_threads=8, float_label_dense=True,
                             check_type=hugectr.Check_t.Sum)
    t0 = time.perf_counter()
    DataGenerator(gp).generate()
    gen_s = time.perf_counter() - t0
    solver = hugectr.CreateSolver(batchsize=batch, batchsize_eval=batch, lr=0.01, vvgpu=[[0]], repeat_dataset=True,
                                  i64_input_key=fmt != "raw", use_cuda_graph=False)
    if fmt == "parquet":
        rp = hugectr.DataReaderParams(kind, source=[src], eval_source=ev, check_type=hugectr.Check_t.Non,
                                      slot_size_array=SLOTS, num_workers=threads)
    elif fmt == "raw":
        rp = hugectr.DataReaderParams(kind, source=[src], eval_source=ev, check_type=hugectr.Check_t.Non,
                                      num_samples=n, eval_num_samples=batch, float_label_dense=True,
                                      slot_size_array=SLOTS,
                                      async_param=hugectr.AsyncParam(threads, 4, 2, 2, 512000, True, hugectr.Alignment_t.Non))
    else:
        rp = hugectr.DataReaderParams(kind, source=[src], eval_source=ev, check_type=hugectr.Check_t.Sum,
                                      slot_size_array=SLOTS, num_workers=threads)
    m = hugectr.Model(solver, rp, hugectr.CreateOptimizer(hugectr.Optimizer_t.SGD))
    m.add(hugectr.Input(label_dim=1, label_name="label", dense_dim=13, dense_name="dense",
                        data_reader_sparse_param_array=[hugectr.DataReaderSparseParam("data1", 1, True, 26)]))
    m.add(hugectr.SparseEmbedding(hugectr.Embedding_t.LocalizedSlotSparseEmbeddingHash, 64, 8, "sum",
                                  "emb", "data1", slot_size_array=SLOTS))
    m.add(hugectr.DenseLayer(hugectr.Layer_t.Reshape, ["emb"], ["r"], leading_dim=26 * 8))
    m.add(hugectr.DenseLayer(hugectr.Layer_t.Concat, ["r", "dense"], ["c"]))
    m.add(hugectr.DenseLayer(hugectr.Layer_t.InnerProduct, ["c"], ["fc"], num_output=1))
    m.add(hugectr.DenseLayer(hugectr.Layer_t.BinaryCrossEntropyLoss, ["fc", "label"], ["loss"]))
    m.compile()
    data = [os.path.join(d, "train")] if fmt == "parquet" else \
        ([l.strip() for l in open(src).read().split()[1:]] if fmt == "norm" else [src])
    return m, dir_bytes(data), gen_s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--formats", default="parquet,raw,norm")
    ap.add_argument("--samples", type=int, default=1 << 20)
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--dir", default="")
    a = ap.parse_args()
    root = a.dir or tempfile.mkdtemp(prefix="hctr_reader_bench_")
    try:
        for fmt in a.formats.split(","):
            m, nbytes, gen_s = build(fmt, root, a.samples, a.batch, a.threads)
            if m.device.type == "cpu":
                torch.set_num_threads(1)   # the consumer copies stand in for H2D DMA: no intra-op threads spinning on the reader cores
            rd = m.get_data_reader_train()
            for _ in range(3):                       # warm-up (thread start, first row groups)
                rd.read_a_batch_to_device()
            if m.device.type == "cuda":
                torch.cuda.synchronize()
            n, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < a.seconds:
                rd.read_a_batch_to_device()
                n += 1
            if m.device.type == "cuda":
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            sps = n * a.batch / dt
            print(json.dumps({"reader": fmt, "device": str(m.device), "samples_per_s": round(sps),
                              "file_MB_per_s": round(sps * nbytes / a.samples / 1e6, 1),
                              "batch": a.batch, "batches": n, "threads": a.threads,
                              "dataset_samples": a.samples, "dataset_MB": round(nbytes / 1e6, 1),
                              "generate_s": round(gen_s, 2), "host_cores": os.cpu_count()}), flush=True)
            m.close()
    finally:
        if not a.dir:
            shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
