"""Embedding-collection benchmark harness (the role of the reference's benchmarks/embedding_collection:
`benchmark.sh` + `hugectr/train.py`): a DLRM-DCNv2-shaped network over one of the reference's synthetic
table sets, planned by `tools.planner.plan_tables`, timed on the device (CUDA events, max over ranks).

Table sets (`benchmark.sh:36-95`): per group (number of tables, vocabulary, hotness, width).
Ablation switches read by the model itself (`Model._step_body`, reference model_pipeline.cpp:118-286):
SKIP_EMBEDDING, SKIP_BOTTOM_MLP, SKIP_TOP_MLP, SKIP_ALLREDUCE, SKIP_H2D.

    python benchmarks/embedding_collection/benchmark.py --workload 180table_70B_hotness80 --batch_per_gpu 8192
    torchrun --nproc-per-node 8 ... benchmark.py --workload dcnv2
    SKIP_EMBEDDING=1 python ... benchmark.py        # dense-only step time
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from hugectr_b200.models.dlrm import (CRITEO_TB_MULTI_HOT, CRITEO_TB_TABLE_SIZES,  # noqa: E402
                                      build_dlrm_dcnv2)
from hugectr_b200.parallel.comm import Comm  # noqa: E402
from hugectr_b200.tools.planner import plan_tables  # noqa: E402

# name -> (num_table[], vocabulary[], hotness[], ev_size[])
WORKLOADS = {
    "dcnv2": ([1] * 26, CRITEO_TB_TABLE_SIZES, CRITEO_TB_MULTI_HOT, [128] * 26),
    "180table_70B_hotness80": (
        [5, 5, 5, 5, 20, 30, 10, 20, 10, 10, 10, 5, 40, 1, 1],
        [10000, 4000000, 4000000, 50000000, 1000, 10000, 5000000, 4000000, 10, 1000, 10000, 100000, 4000000,
         50000000, 500000000],
        [100, 50, 30, 50, 50, 30, 20, 20, 100, 10, 100, 100, 200, 100, 100],
        [128, 64, 64, 32, 128, 128, 256, 128, 128, 64, 128, 64, 64, 128, 32]),
    "7table_470B_hotness20": (
        [1] * 7, [10000000, 400000000, 1000000000, 5000000000, 1000000000, 10000000, 10000000],
        [80, 20, 20, 40, 1, 1, 1], [256, 64, 128, 32, 128, 64, 128]),
    "510table_110B_hotness5": (
        [100, 150, 20, 50, 150, 20, 20], [1000, 100000, 1000000, 2000000, 4000000, 4000000, 4000000],
        [1, 1, 1, 1, 1, 10, 100], [128] * 7),
    "200table_100B_hotness20": (
        [10, 10, 10, 10, 20, 10, 10, 10, 10, 10, 10, 20, 20, 10, 10, 10, 10],
        [100, 1000, 1000, 10000, 10000, 10000, 100000, 1000000, 2000000, 2000000, 4000000, 4000000, 2000000,
         4000000, 4000000, 4000000, 50000000],
        [1, 1, 5, 20, 100, 1, 1, 1, 1, 1, 1, 1, 10, 20, 30, 50, 100],
        [128] * 12 + [64] + [128] * 4),
}


def expand(spec, cap_rows=0):
    nt, vs, hot, ev = spec
    S = [v for n, v in zip(nt, vs) for _ in range(n)]
    H = [v for n, v in zip(nt, hot) for _ in range(n)]
    E = [v for n, v in zip(nt, ev) for _ in range(n)]
    if cap_rows > 0:
        S = [min(s, cap_rows) for s in S]
    return S, H, E


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="dcnv2", choices=list(WORKLOADS))
    ap.add_argument("--batch_per_gpu", type=int, default=6912)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cap_rows", type=int, default=0, help="cap every vocabulary (small-memory runs)")
    ap.add_argument("--fp32", action="store_true")
    args = ap.parse_args(argv)
    comm = Comm.init_from_env()
    n = comm.world_size
    S, H, E = expand(WORKLOADS[args.workload], args.cap_rows)
    sm, st, rep = plan_tables(S, H, E, n, global_batch=args.batch_per_gpu * n)
    m = build_dlrm_dcnv2(batchsize=args.batch_per_gpu * n, num_gpus=n, table_sizes=S, multi_hot=H,
                         ev_size=E, shard_plan=(sm, st), comm=comm, mixed=not args.fp32)
    m.compile()
    cuda = comm.device.type == "cuda"
    for _ in range(args.warmup):
        m.train()
    if cuda:
        torch.cuda.synchronize()
        comm.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    for _ in range(args.iters):
        m.train()
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device=comm.device)
        if n > 1:
            comm.all_reduce_max(ms) if hasattr(comm, "all_reduce_max") else torch.distributed.all_reduce(
                ms, op=torch.distributed.ReduceOp.MAX)
        ms = float(ms)
    else:
        ms = (time.perf_counter() - t0) / args.iters * 1e3
    out = {"workload": args.workload, "gpus": n, "tables": len(S), "ms_per_iter": ms,
           "samples_per_s": args.batch_per_gpu * n / (ms / 1e3),
           "plan_imbalance": round(rep["imbalance"], 3),
           "skip": {k: os.environ[k] for k in ("SKIP_EMBEDDING", "SKIP_BOTTOM_MLP", "SKIP_TOP_MLP", "SKIP_ALLREDUCE",
                                               "SKIP_H2D") if os.environ.get(k, "0") not in ("0", "")}}
    if comm.rank == 0:
        print(json.dumps(out))
    return out


if __name__ == "__main__":
    main()
    os._exit(0)
