"""Emulate ONE rank of an N-GPU job on a single GPU to tune the owner-side embedding kernels cheaply.

The embedding collection is built for (world=N, rank=r) over a stub communicator in collective mode:
the kernels then run with N source-rank key / gradient buffers exactly as on the real machine (same
lookup descriptors, same filtering of row-sharded tables, same index / reduce+update work) -- only the
buffers are local instead of peer mapped.  Usage:

    python tools_dev/emb_rank_emulator.py --world 8 --rank 0 [--cap-rows 4000000] [--batch 6912]

(HCTR_SHARD_SPLIT=1 in the environment A/Bs the requester-side split.)
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hugectr_b200.data.batch import power_law_keys  # noqa: E402
from hugectr_b200.embedding.collection import (EmbeddingCollection, EmbeddingCollectionConfig,  # noqa: E402
                                               EmbeddingTableConfig)
from hugectr_b200.enums import Optimizer_t  # noqa: E402
from hugectr_b200.models.dlrm import CRITEO_TB_MULTI_HOT, CRITEO_TB_TABLE_SIZES  # noqa: E402
from hugectr_b200.solver import CreateOptimizer  # noqa: E402
from hugectr_b200.tools.planner import generate_plan  # noqa: E402


class EmuComm:
    """rank r of N without peers: collectives are no-ops (the buffers are pre-filled by the tool)"""

    def __init__(self, rank, world, device):
        self.rank, self.world_size, self.device = rank, world, device
        self.p2p_available = False
        self.num_nodes = 1

    def all_gather(self, out, inp): pass
    def all_to_all(self, out, inp): out.copy_(inp)
    def all_reduce(self, t): return t
    def barrier(self): pass
    def broadcast(self, t, src=0): return t
    def all_gather_object(self, obj): return [obj] * self.world_size


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--batch", type=int, default=6912)
    ap.add_argument("--cap-rows", type=int, default=0)
    ap.add_argument("--ev", type=int, default=128)
    ap.add_argument("--cpu", action="store_true", help="reference path smoke test (tiny sizes)")
    a = ap.parse_args()
    dev = torch.device("cpu" if a.cpu else "cuda")
    sizes = list(CRITEO_TB_TABLE_SIZES)
    if a.cap_rows:
        sizes = [min(s, a.cap_rows) for s in sizes]
    hot = list(CRITEO_TB_MULTI_HOT)
    n, N, b = len(sizes), a.world, a.batch
    plan = generate_plan(sizes, hot, N, ev_size=a.ev)
    cfg = EmbeddingCollectionConfig()
    cfg.embedding_lookup([EmbeddingTableConfig(str(i), sizes[i], a.ev) for i in range(n)],
                         [f"d{i}" for i in range(n)], "emb", ["sum"] * n)
    cfg.shard(plan[0], plan[1])
    act = torch.float32 if a.cpu else torch.bfloat16
    e = EmbeddingCollection(cfg, b, {f"d{i}": hot[i] for i in range(n)}, dev, act, EmuComm(a.rank, N, dev),
                            CreateOptimizer(Optimizer_t.AdaGrad, epsilon=1e-8), fused=False)
    g = torch.Generator().manual_seed(0)
    for r in range(N):                                   # every source rank's batch
        keys = torch.cat([power_law_keys(b * hot[i], sizes[i], 1.1, g).to(e.key_dtype) for i in range(n)])
        e.keys_all[r, :keys.numel()].copy_(keys)
    e.key_slab.copy_(e.keys_all[a.rank])
    if e.nnz_slab is not None:                           # requester-side split of every source batch
        from hugectr_b200.embedding import ops as E
        for r in range(N):
            tmp_k, tmp_n = e.keys_all[r], e.nnz_all[r]
            for gl in e.glookups:
                if "split_off" in gl:
                    E.shard_split(tmp_k, gl["key_off"], b, gl["hotness"], gl["k"], gl["split_off"], tmp_n,
                                  gl["split_nnz_off"])
    e.grads_all.normal_(0, 0.01)
    lr = torch.tensor([0.004], device=dev)
    st = torch.ones(1, dtype=torch.int32, device=dev)
    mp = [grp for grp in e.groups if grp.kind == "mp"]

    def timeit(fn, reps=5):
        fn()
        if dev.type == "cuda":
            torch.cuda.synchronize()
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(reps):
                fn()
            t.record()
            torch.cuda.synchronize()
            return s.elapsed_time(t) / reps * 1e3
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e6

    def bwd():
        e.backward_index()
        e._index_done = False
        for grp in mp:
            kb, gb = e._bwd_bufs(grp)
            e._accum_update(grp, kb, gb, lr, st)

    print(f"world {N} rank {a.rank}: local mp lookups {sum(len(g_.lookups) for g_ in mp)}, "
          f"split={'on' if e.nnz_slab is not None else 'off'}")
    print("forward (owner side, all source ranks)  us:", round(timeit(e.forward_compute), 1))
    print("backward index + reduce + update        us:", round(timeit(bwd), 1))


if __name__ == "__main__":
    main()
