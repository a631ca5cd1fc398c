"""Embedding-collection benchmark harness (benchmarks/embedding_collection): synthetic workloads
(`dcnv2`, `7table_470B_hotness20`, `180table_70B_hotness80`, ...) with the reference's ablation
switches SKIP_EMBEDDING / SKIP_BOTTOM_MLP / SKIP_TOP_MLP / SKIP_ALLREDUCE / SKIP_H2D
(benchmarks/embedding_collection/README.md:21-29) to attribute step time."""
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
from hugectr_b200.tools.planner import generate_plan  # noqa: E402

WORKLOADS = {
    "dcnv2": (CRITEO_TB_TABLE_SIZES, CRITEO_TB_MULTI_HOT, 128),
    "7table_470B_hotness20": ([int(470e9 / 7 / 128 / 4 / 64)] * 7, [20] * 7, 128),
    "180table_70B_hotness80": ([int(70e9 / 180 / 128 / 4 / 16)] * 180, [80] * 180, 128),
    "200table_100B_hotness20": ([int(100e9 / 200 / 128 / 4 / 16)] * 200, [20] * 200, 128),
    "510table_110B_hotness5": ([int(110e9 / 510 / 128 / 4 / 16)] * 510, [5] * 510, 128),
}

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="dcnv2", choices=list(WORKLOADS))
ap.add_argument("--batch_per_gpu", type=int, default=6912)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
comm = Comm.init_from_env()
n = comm.world_size
tables, hot, ev = WORKLOADS[args.workload]
plan = generate_plan(tables, hot, n, ev_size=ev)
m = build_dlrm_dcnv2(batchsize=args.batch_per_gpu * n, num_gpus=n, table_sizes=tables, multi_hot=hot,
                     ev_size=ev, shard_plan=plan, comm=comm)
m.compile()
skip_emb = os.environ.get("SKIP_EMBEDDING", "0") == "1"
skip_bottom = os.environ.get("SKIP_BOTTOM_MLP", "0") == "1"
skip_top = os.environ.get("SKIP_TOP_MLP", "0") == "1"
if skip_emb:
    m.freeze_embedding()
    for e in m.ebcs_train:
        e.forward_compute = lambda: None
if skip_bottom:
    m.net_train.bottom_layers = []
if skip_top:
    m.freeze_dense()
if os.environ.get("SKIP_ALLREDUCE", "0") == "1":
    m.exchange_wgrad.allreduce = lambda: None
for _ in range(5):
    m.train()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(args.iters):
    if os.environ.get("SKIP_H2D", "0") == "1":
        m._run_step()
    else:
        m.train()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / args.iters
if comm.rank == 0:
    print(json.dumps({"workload": args.workload, "gpus": n, "ms_per_iter": dt * 1e3,
                      "samples_per_s": args.batch_per_gpu * n / dt}))
os._exit(0)
