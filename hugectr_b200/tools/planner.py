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
                  bytes_per_elem: int = 8, num_nodes: int = 1) -> Tuple[List[List[int]], list]:
    """``num_nodes`` > 1 (hier-auto of the reference planner): the shards of a row-split table are
    kept inside ONE node (the least loaded) so that the hierarchical exchange reduces their partial
    sums before anything crosses the node boundary."""
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
    gpn = num_gpus // max(1, num_nodes)
    for t in sorted(mp, key=lambda t: -cost(t, shards[t]) * shards[t]):
        k = shards[t]
        cand = range(num_gpus)
        if num_nodes > 1 and 1 < k <= gpn:
            node_load = [sum(load[nd * gpn:(nd + 1) * gpn]) for nd in range(num_nodes)]
            nd = min(range(num_nodes), key=lambda i: node_load[i])
            cand = range(nd * gpn, (nd + 1) * gpn)
        order = sorted(cand, key=lambda g: (load[g], mem[g]))
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


def plan_report(table_sizes: Sequence[int], multi_hot: Sequence[int], shard_matrix, ev_size: int = 128,
                bytes_per_elem: int = 8) -> dict:
    """per-GPU lookup load (rows gathered per sample) and table memory (GB) of a plan"""
    g = len(shard_matrix)
    n = len(table_sizes)
    k = [sum(shard_matrix[r][t] for r in range(g)) for t in range(n)]
    load = [sum(multi_hot[t] / k[t] for t in range(n) if shard_matrix[r][t]) for r in range(g)]
    mem = [sum(table_sizes[t] * ev_size * bytes_per_elem / 1e9 / k[t] for t in range(n)
               if shard_matrix[r][t]) for r in range(g)]
    return {"shards_per_table": k, "lookups_per_sample": load, "memory_gb": mem,
            "imbalance": max(load) / (sum(load) / g) if sum(load) else 1.0}


def save_plan(path: str, shard_matrix, shard_strategy, column_wise=None):
    import json
    with open(path, "w") as f:
        json.dump({"shard_matrix": shard_matrix, "shard_strategy": [[k, list(v)] for k, v in shard_strategy],
                   "shard_column_wise": column_wise or []}, f, indent=1)


def load_plan(path: str):
    import json
    with open(path) as f:
        d = json.load(f)
    return d["shard_matrix"], [(k, v) for k, v in d["shard_strategy"]], d.get("shard_column_wise", [])


def main(argv=None):
    """python -m hugectr_b200.tools.planner --num-gpus 8 [--num-nodes 1] [--plan auto] --out plan.json
    (Criteo-TB DLRM-DCNv2 tables by default; --slot-sizes / --multi-hot override)"""
    import argparse
    from ..models.dlrm import CRITEO_TB_MULTI_HOT, CRITEO_TB_TABLE_SIZES
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-gpus", type=int, required=True)
    ap.add_argument("--num-nodes", type=int, default=1)
    ap.add_argument("--plan", default="auto", choices=["auto", "round_robin", "uniform"])
    ap.add_argument("--slot-sizes", default="")
    ap.add_argument("--multi-hot", default="")
    ap.add_argument("--ev-size", type=int, default=128)
    ap.add_argument("--out", default="")
    a = ap.parse_args(argv)
    sizes = [int(x) for x in a.slot_sizes.split(",")] if a.slot_sizes else CRITEO_TB_TABLE_SIZES
    hot = [int(x) for x in a.multi_hot.split(",")] if a.multi_hot else CRITEO_TB_MULTI_HOT
    sm, st = generate_plan(sizes, hot, a.num_gpus, ev_size=a.ev_size, plan=a.plan, num_nodes=a.num_nodes)
    rep = plan_report(sizes, hot, sm, a.ev_size)
    print("lookups/sample per GPU:", [round(x, 1) for x in rep["lookups_per_sample"]])
    print("table memory per GPU (GB):", [round(x, 1) for x in rep["memory_gb"]])
    print("imbalance (max/mean): %.3f" % rep["imbalance"])
    if a.out:
        save_plan(a.out, sm, st)
    return sm, st


if __name__ == "__main__":
    main()
