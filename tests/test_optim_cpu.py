"""Dense / sparse optimizers, regularizers and ranking metrics against independent oracles
(torch.optim, closed forms, scikit-learn) -- the pattern of test/utest/optimizer/optimizer_test.cpp,
regularizers/*, metrics/*.cpp of the reference."""
import math

import numpy as np
import pytest
import torch

import hugectr_b200 as hugectr
from hugectr_b200.embedding import ops as E
from hugectr_b200.ops import dense as D


def _run_dense(code, hp, steps, lr, n=257, need=1):
    torch.manual_seed(0)
    w = torch.randn(n)
    w0 = w.clone()
    grads = [torch.randn(n) * 0.3 for _ in range(steps)]
    s0 = torch.zeros(n) if need >= 1 else None
    s1 = torch.zeros(n) if need >= 2 else None
    lr_t = torch.tensor([lr])
    for t, g in enumerate(grads, 1):
        D.dense_opt_step(code, w, g.clone(), None, s0, s1, lr_t, torch.tensor([t], dtype=torch.int32), hp)
    return w0, grads, w


def _torch_ref(make_opt, w0, grads):
    p = torch.nn.Parameter(w0.clone())
    opt = make_opt([p])
    for g in grads:
        p.grad = g.clone()
        opt.step()
    return p.detach()


def test_dense_sgd_momentum_nesterov_adam_adagrad_rmsprop_match_torch():
    lr = 0.05
    w0, g, w = _run_dense(D.D_SGD, {}, 5, lr, need=0)
    torch.testing.assert_close(w, _torch_ref(lambda p: torch.optim.SGD(p, lr=lr), w0, g))
    # HugeCTR momentum: v = mu v - lr g ; w += v   == torch SGD momentum with lr folded in
    w0, g, w = _run_dense(D.D_MOMENTUM, {"momentum": 0.9}, 6, lr)
    torch.testing.assert_close(w, _torch_ref(lambda p: torch.optim.SGD(p, lr=lr, momentum=0.9), w0, g),
                               atol=1e-5, rtol=1e-5)
    w0, g, w = _run_dense(D.D_NESTEROV, {"momentum": 0.9}, 6, lr)
    torch.testing.assert_close(w, _torch_ref(lambda p: torch.optim.SGD(p, lr=lr, momentum=0.9, nesterov=True),
                                             w0, g), atol=1e-5, rtol=1e-5)
    # Adam: epsilon sits outside the bias-corrected sqrt in both formulations up to O(eps)
    w0, g, w = _run_dense(D.D_ADAM, {"beta1": 0.9, "beta2": 0.999, "epsilon": 1e-8}, 8, 0.01, need=2)
    torch.testing.assert_close(w, _torch_ref(lambda p: torch.optim.Adam(p, lr=0.01, eps=1e-8), w0, g),
                               atol=1e-5, rtol=1e-4)
    w0, g, w = _run_dense(D.D_ADAGRAD, {"epsilon": 1e-10}, 6, 0.1)
    torch.testing.assert_close(w, _torch_ref(lambda p: torch.optim.Adagrad(p, lr=0.1, eps=1e-10), w0, g),
                               atol=1e-5, rtol=1e-4)
    w0, g, w = _run_dense(D.D_RMSPROP, {"beta2": 0.95, "epsilon": 1e-8}, 6, 0.01)
    torch.testing.assert_close(w, _torch_ref(lambda p: torch.optim.RMSprop(p, lr=0.01, alpha=0.95, eps=1e-8),
                                             w0, g), atol=1e-5, rtol=1e-4)


def test_dense_ftrl_closed_form_and_loss_scaler():
    # FTRL-proximal (McMahan et al.), one coordinate followed by hand
    lr, l1, l2, beta = 0.1, 0.01, 0.1, 1.0
    w = torch.tensor([0.3])
    z, n = torch.zeros(1), torch.zeros(1)
    zz = nn = 0.0
    ww = 0.3
    for t, g in enumerate([0.5, -0.2, 0.1], 1):
        D.dense_opt_step(D.D_FTRL, w, torch.tensor([g]), None, z, n, torch.tensor([lr]),
                         torch.tensor([t], dtype=torch.int32),
                         {"lambda1": l1, "lambda2": l2, "ftrl_beta": beta})
        sigma = (math.sqrt(nn + g * g + beta) - math.sqrt(nn + beta)) / lr
        zz += g - sigma * ww
        nn += g * g
        ww = 0.0 if abs(zz) <= l1 else (math.copysign(l1, zz) - zz) / (math.sqrt(nn + beta) / lr + l2)
        assert abs(w.item() - ww) < 1e-6, (t, w.item(), ww)
    # loss scaler: gradients arrive multiplied by `scaler` and are divided inside the optimizer
    w1, w2 = torch.ones(8), torch.ones(8)
    g = torch.randn(8)
    one = torch.tensor([1], dtype=torch.int32)
    D.dense_opt_step(D.D_SGD, w1, g.clone(), None, None, None, torch.tensor([0.1]), one, {"scaler": 1.0})
    D.dense_opt_step(D.D_SGD, w2, g * 1024, None, None, None, torch.tensor([0.1]), one, {"scaler": 1024.0})
    torch.testing.assert_close(w1, w2)


@pytest.mark.parametrize("opt,mk,hp", [
    (hugectr.Optimizer_t.SGD, lambda p: torch.optim.SGD(p, lr=0.1), {}),
    (hugectr.Optimizer_t.AdaGrad, lambda p: torch.optim.Adagrad(p, lr=0.1, eps=1e-7), {"epsilon": 1e-7}),
    (hugectr.Optimizer_t.Adam, lambda p: torch.optim.Adam(p, lr=0.1, eps=1e-7), {"epsilon": 1e-7}),
])
def test_sparse_row_optimizers_match_torch(opt, mk, hp):
    """the per-row rule applied to the touched rows == a dense torch optimizer restricted to them"""
    torch.manual_seed(2)
    w = torch.randn(5, 4)
    p = torch.nn.Parameter(w.clone())
    o = mk([p])
    s0, s1 = torch.zeros(5, 4), torch.zeros(5, 4)
    for step in range(1, 5):
        g = torch.randn(5, 4)
        E.sparse_opt_reference(opt, w, s0, s1, g.clone(), dict(hp, beta1=0.9, beta2=0.999), 0.1, step)
        p.grad = g.clone()
        o.step()
    torch.testing.assert_close(w, p.detach(), atol=1e-5, rtol=1e-4)


def test_regularizers_add_penalty_and_gradient():
    from hugectr_b200.layers.base import ParamArena
    from hugectr_b200.layers.loss import Regularizer
    arena = ParamArena()
    p = arena.add("w", (3, 4), lambda shape, gen: torch.randn(shape, generator=gen))
    arena.finalize(torch.device("cpu"), False)
    arena.init_params(1)
    w = p.w.clone()
    lam, batch = 0.02, 16
    for kind, pen, grad in ((hugectr.Regularizer_t.L2, lambda w: 0.5 * lam * (w * w).sum() / batch,
                             lambda w: lam * w / batch),
                            (hugectr.Regularizer_t.L1, lambda w: lam * w.abs().sum() / batch,
                             lambda w: lam * torch.sign(w) / batch)):
        r = Regularizer(kind, lam, [p], batch)
        p.g.zero_()
        r.init_wgrad()
        torch.testing.assert_close(r.rterm().reshape(()), pen(w), atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(p.g, grad(w), atol=1e-7, rtol=1e-5)


def test_ranking_metrics_against_closed_forms():
    from hugectr_b200 import metrics as M
    from hugectr_b200.enums import MetricsRawType as R
    from sklearn.metrics import ndcg_score
    torch.manual_seed(4)
    pred, label = torch.rand(64, 1), (torch.rand(64, 1) > 0.6).float()
    raw = {R.Pred: pred, R.Label: label, R.Loss: torch.tensor([0.7])}
    hr = M.create_metric(hugectr.MetricsType.HitRate, None)
    hr.local_reduce(raw)
    chk = pred > 0.8
    assert abs(hr.finalize_metric() - (chk & (label == 1)).sum().item() / chk.sum().item()) < 1e-6
    nd = M.create_metric(hugectr.MetricsType.NDCG, None)
    nd.local_reduce(raw)
    assert abs(nd.finalize_metric() -
               ndcg_score(label.reshape(1, -1).numpy(), pred.reshape(1, -1).numpy())) < 1e-5
    sm = M.create_metric(hugectr.MetricsType.SMAPE, None)
    y = torch.rand(64, 1) + 0.5
    sm.local_reduce({R.Pred: pred, R.Label: y})
    assert abs(sm.finalize_metric() - ((pred - y).abs() / ((pred + y) / 2)).mean().item()) < 1e-6
    al = M.create_metric(hugectr.MetricsType.AverageLoss, None)
    for v in (0.5, 0.7, 0.9):
        al.local_reduce({R.Loss: torch.tensor([v])})
    assert abs(al.finalize_metric() - 0.7) < 1e-6
    # multi-class AUC = macro average of the per-class AUCs
    from sklearn.metrics import roc_auc_score
    p3, y3 = torch.rand(200, 3), (torch.rand(200, 3) > 0.5).float()
    auc = M.create_metric(hugectr.MetricsType.AUC, None, num_classes=3)
    auc.local_reduce({R.Pred: p3, R.Label: y3})
    exp = np.mean([roc_auc_score(y3[:, c].numpy(), p3[:, c].numpy()) for c in range(3)])
    assert abs(auc.finalize_metric() - exp) < 1e-6


@pytest.mark.parametrize("kind", ["adam", "momentum", "nesterov"])
def test_legacy_embedding_global_update_equals_dense_optimizer(kind):
    """Update_t.Global on the legacy hashed embedding == the dense torch optimizer over the whole
    table (untouched rows keep decaying / moving), sparse_optimizer.cu:241-292"""
    from hugectr_b200.data.batch import HostBatch
    from hugectr_b200.embedding.sparse_embedding import SparseEmbeddingRuntime
    from hugectr_b200.parallel.comm import Comm
    torch.manual_seed(0)
    b, S, H, vec, vocab = 8, 2, 2, 4, 40
    opt = {"adam": hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam, hugectr.Update_t.Global, epsilon=1e-7),
           "momentum": hugectr.CreateOptimizer(hugectr.Optimizer_t.MomentumSGD, hugectr.Update_t.Global,
                                               momentum_factor=0.9),
           "nesterov": hugectr.CreateOptimizer(hugectr.Optimizer_t.Nesterov, hugectr.Update_t.Global,
                                               momentum_factor=0.9)}[kind]
    cfg = hugectr.SparseEmbedding(hugectr.Embedding_t.DistributedSlotSparseEmbeddingHash, 0, vec, "sum",
                                  "emb", "data", slot_size_array=[vocab // 2] * 2, optimizer=opt)
    cfg.max_vocabulary_size_per_gpu = 64
    prm = hugectr.DataReaderSparseParam("data", H, True, S)
    dev = torch.device("cpu")
    rt = SparseEmbeddingRuntime(cfg, prm, None, b, dev, torch.float32, Comm.single(dev), opt, torch.int64)
    W = torch.nn.Parameter(rt.table.view(-1, vec).clone())
    lr = 0.05
    # Adam: the framework's own dense rule (epsilon outside the bias-corrected root, Appendix A.2) --
    # checked against torch.optim.Adam above; torch's epsilon placement differs for near-zero moments
    m_d, v_d = torch.zeros_like(W), torch.zeros_like(W)
    topt = {"adam": lambda: None,
            "momentum": lambda: torch.optim.SGD([W], lr=lr, momentum=0.9),
            "nesterov": lambda: torch.optim.SGD([W], lr=lr, momentum=0.9, nesterov=True)}[kind]()
    for step in range(1, 6):
        keys = torch.randint(0, vocab, (b * S * H,))
        hb = HostBatch(torch.zeros(b, 1), torch.zeros(b, 0), keys, None, b)
        rt.set_keys(hb, {"data": 0}, {})
        rt.forward(True)
        g = torch.randn(b, S, vec)
        rt.top_grad.copy_(g)
        rows = rt.rows_all.view(b, S, H).clone()
        rt.backward(torch.tensor([lr]), torch.tensor([step], dtype=torch.int32))
        dense = torch.zeros_like(W)
        dense.index_add_(0, rows.reshape(-1), g.unsqueeze(2).expand(b, S, H, vec).reshape(-1, vec))
        if kind == "adam":
            D.dense_opt_reference(D.D_ADAM, W.data, dense, None, m_d, v_d, lr, step,
                                  {"beta1": 0.9, "beta2": 0.999, "epsilon": 1e-7})
        else:
            W.grad = dense
            topt.step()
    n = rt.hash.size()
    torch.testing.assert_close(rt.table.view(-1, vec)[:n], W.detach()[:n], atol=2e-5, rtol=1e-4)


def test_mean_combiner_averages_over_the_actual_bag():
    """padded (-1) keys: 'mean' divides by the number of valid keys of the bag, forward and backward"""
    from hugectr_b200.embedding.collection import (EmbeddingCollection, EmbeddingCollectionConfig,
                                                   EmbeddingTableConfig)
    from hugectr_b200.parallel.comm import Comm
    dev = torch.device("cpu")
    cfg = EmbeddingCollectionConfig()
    cfg.embedding_lookup([EmbeddingTableConfig("0", 50, 4), EmbeddingTableConfig("1", 20, 4)], ["d0", "d1"],
                         "emb", ["mean", "sum"])
    e = EmbeddingCollection(cfg, 3, {"d0": 4, "d1": 2}, dev, torch.float32, Comm.single(dev),
                            hugectr.CreateOptimizer(hugectr.Optimizer_t.SGD))
    k0 = torch.tensor([[1, 2, -1, -1], [3, 3, 3, 3], [5, -1, -1, -1]])
    k1 = torch.tensor([[1, 2], [3, -1], [4, 4]])
    e.set_keys(torch.cat([k0.reshape(-1), k1.reshape(-1)]).int())
    W = e.groups[0].table.view(-1, 4).clone()
    e.forward()
    off1 = [sl["row_off"] for sl in e.groups[0].table_slices if sl["table"] == "1"][0]
    torch.testing.assert_close(e.top_data["emb"][:, :4], torch.stack([(W[1] + W[2]) / 2, W[3], W[5]]))
    g = torch.randn(3, 8)
    e.top_grad["emb"].copy_(g)
    e.backward(torch.tensor([1.0]), torch.tensor([1], dtype=torch.int32))
    exp = W.clone()
    exp[1] -= g[0, :4] / 2
    exp[2] -= g[0, :4] / 2
    exp[3] -= g[1, :4]
    exp[5] -= g[2, :4]
    exp[off1 + 1] -= g[0, 4:]
    exp[off1 + 2] -= g[0, 4:]
    exp[off1 + 3] -= g[1, 4:]
    exp[off1 + 4] -= 2 * g[2, 4:]
    torch.testing.assert_close(e.groups[0].table.view(-1, 4), exp, atol=1e-6, rtol=1e-5)


def test_legacy_embedding_lazy_global_adam_follows_reference_kernel():
    """Update_t.LazyGlobal: per-row catch-up with the skipped steps, then moment update
    (opt_adam_kernel_lazy, sparse_optimizer.cu:523-561) -- compared with a scalar re-implementation"""
    from hugectr_b200.data.batch import HostBatch
    from hugectr_b200.embedding.sparse_embedding import SparseEmbeddingRuntime
    from hugectr_b200.parallel.comm import Comm
    torch.manual_seed(0)
    b, S, H, vec, vocab = 4, 2, 1, 3, 12
    opt = hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam, hugectr.Update_t.LazyGlobal, epsilon=1e-7)
    cfg = hugectr.SparseEmbedding(hugectr.Embedding_t.DistributedSlotSparseEmbeddingHash, 0, vec, "sum",
                                  "emb", "data", slot_size_array=[6, 6], optimizer=opt)
    cfg.max_vocabulary_size_per_gpu = 32
    dev = torch.device("cpu")
    rt = SparseEmbeddingRuntime(cfg, hugectr.DataReaderSparseParam("data", H, True, S), None, b, dev,
                                torch.float32, Comm.single(dev), opt, torch.int64)
    W = rt.table.view(-1, vec).clone().double()
    M, V = torch.zeros_like(W), torch.zeros_like(W)
    P = torch.ones(W.shape[0], dtype=torch.float64)
    lr, b1, b2, eps = 0.05, 0.9, 0.999, 1e-7
    for step in range(1, 7):
        keys = torch.randint(0, vocab, (b * S * H,))
        rt.set_keys(HostBatch(torch.zeros(b, 1), torch.zeros(b, 0), keys, None, b), {"data": 0}, {})
        rt.forward(True)
        g = torch.randn(b, S, vec)
        rt.top_grad.copy_(g)
        rows = rt.rows_all.view(b, S, H).clone()
        rt.backward(torch.tensor([lr]), torch.tensor([step], dtype=torch.int32))
        gsum = {}
        for i in range(b):
            for s in range(S):
                r = int(rows[i, s, 0])
                gsum[r] = gsum.get(r, 0) + g[i, s].double()
        for r, gi in gsum.items():
            prev = P[r].item()
            skipped = step - prev
            alpha_t = lr / (1 - b1) * (1 - b2 ** prev) ** 0.5 / (1 - b1 ** prev) * (1 - b1 ** skipped)
            W[r] += -alpha_t * M[r] / (V[r].sqrt() + eps)
            M[r] = b1 ** skipped * M[r] + (1 - b1) * gi
            V[r] = b2 ** skipped * V[r] + (1 - b2) * gi * gi
            P[r] = step
    n = rt.hash.size()
    torch.testing.assert_close(rt.table.view(-1, vec)[:n].double(), W[:n], atol=1e-5, rtol=1e-4)
