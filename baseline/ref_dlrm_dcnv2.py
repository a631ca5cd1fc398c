#!/usr/bin/env python
"""Reference arm of bench.py: the UNMODIFIED reference (`baseline/_ref/hugectr.so`, built from /root/reference
by `baseline/build_reference.py`) training DLRM-DCNv2 through its own public Python API and stock code path
(RawAsync multi-hot reader -> EmbeddingCollection -> MLP / MultiCross -> AdaGrad), the model of
/root/reference/samples/dlrm/train.py:300-500 with the same solver switches (mixed precision, CUDA graph,
intra/inter-iteration overlap, grouped all-reduce).  Nothing of hugectr_b200 is imported here.

One process drives all N GPUs (the reference's execution model).  Data: a synthetic Criteo-TB shaped raw file
(label int32, 13 float dense, 214 int32 keys per sample; power-law keys alpha=1.1) written to local disk first.
Timing: wall clock around K `model.train()` calls bracketed by `get_current_loss()` (a blocking D2H of the loss
= device sync); that is the end-to-end number (file -> pinned host -> H2D -> step), and also the only number
the reference's API exposes, so `value` == `e2e.value`.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TABLES = [40000000, 39060, 17295, 7424, 20265, 3, 7122, 1543, 63, 40000000, 3067956, 405282, 10, 2209, 11938,
          155, 4, 976, 14, 40000000, 40000000, 40000000, 590152, 12973, 108, 36]
HOT = [3, 2, 1, 2, 6, 1, 1, 1, 1, 7, 3, 8, 1, 6, 9, 5, 1, 1, 1, 12, 100, 27, 10, 3, 1, 1]
NUM_DENSE = 13


def power_law(rng, n, vocab, alpha=1.1):
    """inverse-CDF sampler, same law as the reference DataGenerator (data_generator.hpp:109-131)"""
    if vocab <= 1:
        return np.zeros(n, dtype=np.int32)
    u = rng.random(n)
    a = 1.0 - alpha
    x = ((float(vocab + 1) ** a - 1.0) * u + 1.0) ** (1.0 / a)
    return np.clip(np.floor(x).astype(np.int64) - 1, 0, vocab - 1).astype(np.int32)


def write_raw(path, num_samples, tables, seed=1):
    rng = np.random.default_rng(seed)
    cols = 1 + NUM_DENSE + sum(HOT)
    chunk = 1 << 16
    with open(path, "wb") as f:
        done = 0
        while done < num_samples:
            n = min(chunk, num_samples - done)
            rec = np.empty((n, cols), dtype=np.int32)
            rec[:, 0] = (rng.random(n) < 0.3).astype(np.int32)
            rec[:, 1:1 + NUM_DENSE] = rng.random((n, NUM_DENSE), dtype=np.float32).view(np.int32)
            c = 1 + NUM_DENSE
            for t, h in zip(tables, HOT):
                rec[:, c:c + h] = power_law(rng, n * h, t).reshape(n, h)
                c += h
            f.write(rec.tobytes())
            done += n
    return cols * 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--per-gpu-batch", type=int, default=6912)
    ap.add_argument("--table-cap", type=int, default=0)
    ap.add_argument("--plan", default="", help="json file: {shard_matrix:[[names]], shard_strategy:[[kind,[names]]]}")
    ap.add_argument("--file-batches", type=int, default=24)
    ap.add_argument("--data-dir", default="/tmp/hctr_ref_data")
    ap.add_argument("--optimizer", default="adagrad")
    args = ap.parse_args()

    sys.path.insert(0, os.path.join(HERE, "_ref"))
    import hugectr                                       # the reference's pybind11 module
    assert hugectr.__file__.startswith(os.path.join(HERE, "_ref")), hugectr.__file__

    n = args.gpus
    gb = args.per_gpu_batch * n
    tables = [min(t, args.table_cap) if args.table_cap > 0 else t for t in TABLES]
    os.makedirs(args.data_dir, exist_ok=True)
    num_samples = gb * args.file_batches
    path = os.path.join(args.data_dir, f"train_{num_samples}_{args.table_cap}.bin")
    t0 = time.time()
    if not os.path.exists(path) or os.path.getsize(path) != num_samples * 912:
        write_raw(path, num_samples, tables)
    t_data = time.time() - t0

    if args.plan:
        pl = json.load(open(args.plan))
        shard_matrix = pl["shard_matrix"]
        shard_strategy = [(k, [str(x) if not isinstance(x, list) else (str(x[0]), int(x[1])) for x in v])
                          for k, v in pl["shard_strategy"]]
    else:   # round robin, every table model-parallel on one GPU
        shard_matrix = [[] for _ in range(n)]
        for i in range(len(tables)):
            shard_matrix[i % n].append(str(i))
        shard_strategy = [("mp", [str(i) for i in range(len(tables))])]

    solver = hugectr.CreateSolver(
        model_name="dlrm_dcnv2", seed=0, max_eval_batches=1, batchsize_eval=gb, batchsize=gb,
        vvgpu=[list(range(n))], repeat_dataset=True, lr=0.004, warmup_steps=1, decay_start=0, decay_steps=1,
        decay_power=2.0, end_lr=0.0, use_mixed_precision=True, enable_tf32_compute=False, scaler=1024,
        use_cuda_graph=True, gen_loss_summary=True, train_intra_iteration_overlap=True,
        train_inter_iteration_overlap=True, eval_intra_iteration_overlap=False,
        eval_inter_iteration_overlap=True, all_reduce_algo=hugectr.AllReduceAlgo.NCCL, grouped_all_reduce=True,
        num_iterations_statistics=20, perf_logging=False, drop_incomplete_batch=True,
        use_embedding_collection=True, use_algorithm_search=True)
    if args.optimizer == "adagrad":
        optimizer = hugectr.CreateOptimizer(optimizer_type=hugectr.Optimizer_t.AdaGrad,
                                            update_type=hugectr.Update_t.Global, initial_accu_value=0.0,
                                            epsilon=1e-8)
    else:
        optimizer = hugectr.CreateOptimizer(optimizer_type=hugectr.Optimizer_t.SGD,
                                            update_type=hugectr.Update_t.Local, atomic_update=True)
    reader = hugectr.DataReaderParams(
        data_reader_type=hugectr.DataReaderType_t.RawAsync, source=[path], eval_source=path,
        check_type=hugectr.Check_t.Non, num_samples=num_samples, eval_num_samples=num_samples,
        cache_eval_data=1, slot_size_array=tables,
        async_param=hugectr.AsyncParam(num_threads=1, num_batches_per_thread=16, shuffle=False,
                                       multi_hot_reader=True, is_dense_float=True))
    model = hugectr.Model(solver, reader, optimizer)
    nt = len(tables)
    model.add(hugectr.Input(label_dim=1, label_name="label", dense_dim=NUM_DENSE, dense_name="dense",
                            data_reader_sparse_param_array=[
                                hugectr.DataReaderSparseParam(f"data{i}", HOT[i], True, 1) for i in range(nt)]))
    tabs = [hugectr.EmbeddingTableConfig(name=str(i), max_vocabulary_size=tables[i], ev_size=128)
            for i in range(nt)]
    ebc = hugectr.EmbeddingCollectionConfig(use_exclusive_keys=True,
                                            comm_strategy=hugectr.CommunicationStrategy.Uniform)
    ebc.embedding_lookup(table_config=tabs, bottom_name=[f"data{i}" for i in range(nt)],
                         top_name="sparse_embedding", combiner=["sum"] * nt)
    ebc.shard(shard_matrix=shard_matrix, shard_strategy=shard_strategy)
    model.add(ebc)
    cc = hugectr.DenseLayerComputeConfig(async_wgrad=True, fuse_wb=False)
    model.add(hugectr.DenseLayer(layer_type=hugectr.Layer_t.MLP, bottom_names=["dense"], top_names=["mlp1"],
                                 num_outputs=[512, 256, 128], act_type=hugectr.Activation_t.Relu,
                                 compute_config=cc))
    model.add(hugectr.DenseLayer(layer_type=hugectr.Layer_t.Concat, bottom_names=["sparse_embedding", "mlp1"],
                                 top_names=["concat1"]))
    model.add(hugectr.DenseLayer(layer_type=hugectr.Layer_t.MultiCross, bottom_names=["concat1"],
                                 top_names=["interaction1"], projection_dim=512, num_layers=3,
                                 compute_config=cc))
    R, N = hugectr.Activation_t.Relu, hugectr.Activation_t.Non
    model.add(hugectr.DenseLayer(layer_type=hugectr.Layer_t.MLP, bottom_names=["interaction1"],
                                 top_names=["mlp2"], num_outputs=[1024, 1024, 512, 256, 1],
                                 activations=[R, R, R, R, N], compute_config=cc))
    model.add(hugectr.DenseLayer(layer_type=hugectr.Layer_t.BinaryCrossEntropyLoss,
                                 bottom_names=["mlp2", "label"], top_names=["loss"]))
    model.compile()
    model.start_data_reading()

    W, K = max(args.warmup, 3), args.steps
    for _ in range(W + 2):
        model.train()
    loss0 = model.get_current_loss()            # blocking D2H: device sync
    t0 = time.perf_counter()
    for _ in range(K):
        model.train()
    loss1 = model.get_current_loss()
    dt = time.perf_counter() - t0
    # sustained: at least 2 s of steps
    n_sus, t1 = 0, time.perf_counter()
    while time.perf_counter() - t1 < 2.0:
        for _ in range(10):
            model.train()
        model.get_current_loss()
        n_sus += 10
    dt_sus = time.perf_counter() - t1
    value = gb * K / dt
    out = {
        "metric": "DLRM-DCNv2 Criteo-TB training samples/sec (device-timed, max over ranks)",
        "impl": "reference", "value": value, "unit": "samples/s", "n_gpus": n, "steps": K, "warmup": W,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 mixed precision (the reference has no bf16), fp32 master weights / tables / AdaGrad state",
        "data": "synthetic",
        "config": {"model": "DLRM-DCNv2 (MLPerf v3.1 shape; samples/dlrm/train.py model through the reference API)",
                   "global_batch": gb, "per_gpu_batch": args.per_gpu_batch, "seq_len": None,
                   "parallelism": f"single process x {n} GPUs, dp dense + model-parallel embeddings",
                   "table_row_cap": args.table_cap, "optimizer": args.optimizer,
                   "timing": "wall clock around K model.train() calls between two blocking get_current_loss()",
                   "reader": "RawAsync (libaio) from a local synthetic raw file", "loss": [loss0, loss1],
                   "data_gen_s": round(t_data, 1)},
        "e2e": {"value": value, "unit": "samples/s", "ms_per_step": dt / K * 1e3,
                "h2d_bytes_per_step": gb * 912, "d2h_bytes_per_step": 0,
                "note": "same measurement: file -> pinned -> H2D -> step every iteration; loss read once at the end"},
        "sustained": {"value": gb * n_sus / dt_sus, "steps": n_sus, "seconds": dt_sus},
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
