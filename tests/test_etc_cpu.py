"""Embedding training cache through the public API: W&D whose two tables live on the host parameter server
behind the gpu_cache (hugectr.CreateETC, TrainPSType_t.Cached / Staged) must train exactly like the same
model with in-memory hash tables; checkpoints round-trip through the host server."""
import os

import pytest
import torch

import hugectr_b200 as hugectr
from hugectr_b200.models.legacy import build_wdl


def _train(etc, steps=6, seed_weights=None):
    torch.manual_seed(0)
    m = build_wdl(batchsize=64, wide_slot_sizes=[50, 60], deep_slot_sizes=[40] * 6, workspace_mb=(1, 2), lr=0.01,
                  etc=etc)
    m.compile()
    if seed_weights is not None:
        m.arena.weights.copy_(seed_weights)
        m.arena.sync_shadow()
    return m


@pytest.mark.parametrize("ps", [hugectr.TrainPSType_t.Cached, hugectr.TrainPSType_t.Staged])
def test_wdl_cached_tables_match_in_memory_tables(ps):
    ref = _train(None)
    etc = hugectr.CreateETC(ps_types=[ps, ps], sparse_models=["", ""])
    off = _train(etc, seed_weights=ref.arena.weights)
    from hugectr_b200.embedding.offloaded import CachedSparseEmbeddingRuntime
    assert all(isinstance(rt, CachedSparseEmbeddingRuntime) for rt in off.legacy_train)
    pool = ref.reader_train.pool
    ref.train_on_host_batch(pool[0])          # creates the reference's rows for this batch
    # copy the reference tables (weights + optimizer states) and dense state wholesale, then train both on
    # the same, already seen batch: every row the loss depends on is touched in every step
    for r, o in zip(ref.legacy_train, off.legacy_train):
        keys, rows = r.hash.dump()
        o.ps.push(keys, r.table.view(-1, r.vec)[rows], [s.view(-1, r.vec)[rows] for s in (r.s0, r.s1) if s is not None])
    off.arena.weights.copy_(ref.arena.weights)
    off.arena.sync_shadow()
    off.opt_s0.copy_(ref.opt_s0)
    if ref.opt_s1 is not None:
        off.opt_s1.copy_(ref.opt_s1)
    off.step_t.copy_(ref.step_t)
    seen = pool[0]
    for i in range(4):
        torch.manual_seed(100 + i)            # same dropout masks
        ref.train_on_host_batch(seen)
        torch.manual_seed(100 + i)
        off.train_on_host_batch(seen)
        assert abs(ref.get_current_loss() - off.get_current_loss()) < 1e-5
    assert float((ref.arena.weights - off.arena.weights).abs().max()) < 1e-5
    for r, o in zip(ref.legacy_train, off.legacy_train):
        keys = torch.unique(r.keys_loc[r.keys_loc >= 0]).cpu()     # rows of the repeated batch
        rows = r.hash.get(keys.to(r.device)).cpu()
        w_off, _ = o.ps.pull(keys)
        assert float((w_off - r.table.view(-1, r.vec)[rows].cpu()).abs().max()) < 1e-5
    if ps == hugectr.TrainPSType_t.Cached:
        assert off.legacy_train[1].cache.hits > 0


def test_cached_table_checkpoint_roundtrip(tmp_path):
    etc = hugectr.CreateETC(ps_types=[hugectr.TrainPSType_t.Cached] * 2, sparse_models=["", ""])
    m = _train(etc)
    for hb in m.reader_train.pool[:3]:
        m.train_on_host_batch(hb)
    prefix = str(tmp_path / "wdl")
    m.save_params_to_files(prefix, 3)
    etc2 = hugectr.CreateETC(ps_types=[hugectr.TrainPSType_t.Cached] * 2,
                             sparse_models=[prefix + "0_sparse_3.model", prefix + "1_sparse_3.model"])
    m2 = _train(etc2)
    for a, b in zip(m.legacy_train, m2.legacy_train):
        ka, ra = a.ps.items()
        wa, _ = a.ps.pull(ka)
        wb, _ = b.ps.pull(ka)
        assert float((wa - wb).abs().max()) == 0.0
