"""GPU tests of the auxiliary native kernels (hash table, set-associative LRU cache) and GPU smoke
trainings of the legacy-embedding sample models and the model zoo."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_hashtable_kernels_match_dict():
    from hugectr_b200.embedding.hashtable import HashTable
    dev = torch.device("cuda")
    ht = HashTable(5000, dev)
    g = torch.Generator().manual_seed(1)
    keys = torch.randint(0, 3000, (20000,), generator=g)
    keys[::97] = -1
    rows = ht.get_insert(keys.to(dev)).cpu()
    # same key -> same row, distinct keys -> distinct rows, dense numbering, -1 passes through
    seen = {}
    for k, r in zip(keys.tolist(), rows.tolist()):
        if k < 0:
            assert r == -1
            continue
        assert seen.setdefault(k, r) == r
    assert len(set(seen.values())) == len(seen) == ht.size()
    assert sorted(seen.values()) == list(range(len(seen)))
    q = torch.tensor([keys[1].item(), 10 ** 9 + 7, keys[5].item()])
    got = ht.get(q.to(dev)).cpu().tolist()
    assert got[0] == seen[q[0].item()] and got[1] == -1 and got[2] == seen[q[2].item()]
    dk, dr = ht.dump()
    assert dict(zip(dk.cpu().tolist(), dr.cpu().tolist())) == seen
    ht.set(torch.tensor([123456789], device=dev), torch.tensor([4242], device=dev))
    assert ht.get(torch.tensor([123456789], device=dev)).item() == 4242


def test_gpu_cache_query_replace_update_dump():
    from hugectr_b200.cache import GpuCache
    dev = torch.device("cuda")
    c = GpuCache(4096, 8, dev, ways=64)
    keys = torch.arange(1000, device=dev) * 7 + 3
    vals = torch.randn(1000, 8, device=dev)
    out, mi, mk = c.query(keys)
    assert mi.numel() == 1000 and torch.equal(mk, keys)
    c.replace(keys, vals)
    probe = torch.cat([keys[:100], torch.tensor([999999, 888888], device=dev)])
    out, mi, mk = c.query(probe)
    assert mi.tolist() == [100, 101] and mk.tolist() == [999999, 888888]
    torch.testing.assert_close(out[:100], vals[:100])
    c.update(keys[:10], torch.ones(10, 8, device=dev))
    torch.testing.assert_close(c.query(keys[:10])[0], torch.ones(10, 8, device=dev))
    assert set(c.dump().cpu().tolist()) == set(keys.cpu().tolist())
    # LRU eviction: overfill one cache far beyond capacity, recently used keys survive
    small = GpuCache(64, 4, dev, ways=64)           # a single set
    a = torch.arange(64, device=dev)
    small.replace(a, torch.zeros(64, 4, device=dev))
    small.query(a[:32])                              # touch the first half
    ek, ev = small.replace(torch.arange(1000, 1032, device=dev), torch.ones(32, 4, device=dev),
                           return_evicted=True)
    assert set(ek.cpu().tolist()) == set(range(32, 64))
    assert small.query(a[:32])[1].numel() == 0


def _train_cuda(m, iters=6):
    m.compile()
    losses = []
    for _ in range(iters):
        assert m.train()
        losses.append(m.get_current_loss())
    assert np.isfinite(losses).all(), losses
    m.eval()
    assert all(np.isfinite(v) for _, v in m.get_eval_metrics())


@pytest.mark.parametrize("mixed", [False, True])
def test_legacy_models_train_on_gpu(mixed):
    from hugectr_b200.models import build_dcn, build_deepfm, build_wdl
    slots = [300] * 26
    _train_cuda(build_dcn(batchsize=256, slot_sizes=slots, workspace_mb=4, mixed=mixed, max_eval_batches=1))
    _train_cuda(build_deepfm(batchsize=256, slot_sizes=slots, workspace_mb=4, mixed=mixed, max_eval_batches=1))
    _train_cuda(build_wdl(batchsize=256, wide_slot_sizes=[50, 60], deep_slot_sizes=slots,
                          workspace_mb=(1, 4), mixed=mixed, max_eval_batches=1))


@pytest.mark.parametrize("mixed", [False, True])
def test_model_zoo_trains_on_gpu(mixed):
    from hugectr_b200.models import zoo
    _train_cuda(zoo.build_ncf("neumf", batchsize=256, num_users=500, num_items=400, mixed=mixed, max_eval_batches=1))
    _train_cuda(zoo.build_mmoe(batchsize=256, num_slots=8, vocab=200, ev=16, mixed=mixed, max_eval_batches=1))
    _train_cuda(zoo.build_din(batchsize=128, seq_len=6, item_vocab=300, cate_vocab=40, user_vocab=80, ev=8,
                              mixed=mixed, max_eval_batches=1))
    _train_cuda(zoo.build_bst(batchsize=128, seq_len=5, item_vocab=300, user_vocab=80, ev=32, heads=4,
                              mixed=mixed, max_eval_batches=1))
