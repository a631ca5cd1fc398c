"""Sharding planner: greedy cost model -> (shard_matrix, shard_strategy).

Capabilities of the reference planner (samples/dlrm/sharding/planner.py:22-327,
generate_plan.py:23-131): data-parallel threshold for tiny tables, hot-table split (row-wise over
several GPUs) when one table's lookup cost exceeds the balanced share, memory-cap split, then
greedy longest-processing-time placement by cost = hotness * (1 + comm/mem ratio).
B200 numbers: HBM ~6.5 TB/s measured, NVLink ~770 GB/s per direction -> ratio ~8.5.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def generate_plan(table_sizes: Sequence[int], multi_hot: Sequence[int], num_gpus: int,
                  ev_size: int = 128, plan: str = "auto", dp_threshold_rows: int = 4096,
                  mem_comm_bw_ratio: float = 6.5e12 / 770e9, mem_cap_gb: float = 150.0,
                  bytes_per_elem: int = 8) -> Tuple[List[List[int]], list]:
    n = len(table_sizes)
    names = [str(i) for i in range(n)]
    if num_gpus == 1:
        return [[1] * n], [("mp", names)]
    if plan == "round_robin":
        sm = [[1 if (t % num_gpus) == g else 0 for t in range(n)] for g in range(num_gpus)]
        return sm, [("mp", names)]
    if plan == "uniform":      # every table row-sharded over every GPU
        return [[1] * n for _ in range(num_gpus)], [("mp", names)]
    dp = [t for t in range(n) if table_sizes[t] <= dp_threshold_rows]
    mp = [t for t in range(n) if t not in dp]
    # cost of a table = rows gathered per sample (HBM) + pooled vector sent per shard (NVLink)
    def cost(t, k):
        return multi_hot[t] / k + mem_comm_bw_ratio * 0.5
    total = sum(cost(t, 1) for t in mp)
    share = total / num_gpus
    shards = {}
    for t in mp:
        k = 1
        mem_gb = table_sizes[t] * ev_size * bytes_per_elem / 1e9
        while k < num_gpus and (cost(t, k) > 1.25 * share or mem_gb / k > mem_cap_gb):
            k *= 2
        shards[t] = min(k, num_gpus)
    load = [0.0] * num_gpus
    mem = [0.0] * num_gpus
    sm = [[0] * n for _ in range(num_gpus)]
    for t in sorted(mp, key=lambda t: -cost(t, shards[t]) * shards[t]):
        k = shards[t]
        order = sorted(range(num_gpus), key=lambda g: (load[g], mem[g]))
        for g in order[:k]:
            sm[g][t] = 1
            load[g] += cost(t, k)
            mem[g] += table_sizes[t] * ev_size * bytes_per_elem / 1e9 / k
    for t in dp:
        for g in range(num_gpus):
            sm[g][t] = 1
    strategy = []
    if mp:
        strategy.append(("mp", [names[t] for t in mp]))
    if dp:
        strategy.append(("dp", [names[t] for t in dp]))
    return sm, strategy
