"""CPU smoke trainings of the additional model families (NCF, MMoE, DIN, BST) on synthetic data."""
import numpy as np
import pytest
import torch

from hugectr_b200.models import zoo
from hugectr_b200.parallel.comm import Comm

CPU = lambda: Comm.single(torch.device("cpu"))


def _train(m, iters=6):
    m.compile()
    losses = []
    for _ in range(iters):
        assert m.train()
        losses.append(m.get_current_loss())
    assert np.isfinite(losses).all(), losses
    m.eval()
    return losses


@pytest.mark.parametrize("kind", ["gmf", "mlp", "neumf"])
def test_ncf(kind):
    _train(zoo.build_ncf(kind, batchsize=64, num_users=300, num_items=200, comm=CPU(), max_eval_batches=1))


def test_mmoe_multitask():
    m = zoo.build_mmoe(batchsize=64, num_slots=6, vocab=100, ev=8, expert_dims=(32, 16), tower_dim=8,
                       comm=CPU(), max_eval_batches=1, label_weights=[0.7, 0.3])
    _train(m)
    assert len(m.net_train.loss_layers) == 2


def test_din_target_attention():
    _train(zoo.build_din(batchsize=32, seq_len=5, item_vocab=200, cate_vocab=30, user_vocab=50, ev=6,
                         att_dims=(16, 8), mlp_dims=(24, 12), comm=CPU(), max_eval_batches=1))


def test_bst_transformer():
    _train(zoo.build_bst(batchsize=32, seq_len=4, item_vocab=200, user_vocab=50, ev=16, heads=4,
                         ffn_dim=24, mlp_dims=(32, 16), comm=CPU(), max_eval_batches=1))


@pytest.mark.parametrize("name", ["ncf", "mmoe", "din", "bst"])
def test_graph_json_round_trip(name, tmp_path):
    """graph_to_json -> construct_from_json rebuilds an identical, trainable network"""
    import json
    import hugectr_b200 as hugectr
    b = {"ncf": lambda: zoo.build_ncf("neumf", batchsize=32, num_users=100, num_items=80, comm=CPU(),
                                      max_eval_batches=1),
         "mmoe": lambda: zoo.build_mmoe(batchsize=32, num_slots=4, vocab=50, ev=8, expert_dims=(16, 8),
                                        tower_dim=8, comm=CPU(), max_eval_batches=1),
         "din": lambda: zoo.build_din(batchsize=16, seq_len=3, item_vocab=60, cate_vocab=10, user_vocab=20,
                                      ev=4, att_dims=(8, 4), mlp_dims=(8, 4), comm=CPU(), max_eval_batches=1),
         "bst": lambda: zoo.build_bst(batchsize=16, seq_len=3, item_vocab=60, user_vocab=20, ev=8, heads=2,
                                      ffn_dim=8, mlp_dims=(8, 4), comm=CPU(), max_eval_batches=1)}[name]
    m = b()
    m.compile()
    p = str(tmp_path / f"{name}.json")
    m.graph_to_json(p)
    assert len(json.load(open(p))["layers"]) > 5
    m2 = hugectr.Model(m.solver, m.reader_params, m.opt_params, comm=CPU())
    m2.construct_from_json(p, include_dense_network=True)
    m2.compile()
    assert m2.arena.num_params == m.arena.num_params
    assert [r[0] for r in m2.net_train.summary_rows()] == [r[0] for r in m.net_train.summary_rows()]
    assert m2.train()
    assert np.isfinite(m2.get_current_loss())


@pytest.mark.parametrize("name", ["deepfm", "wdl", "ncf", "mmoe", "din", "bst", "dcnv2"])
def test_onnx_converter_matches_model_predictions(name, tmp_path):
    """hugectr2onnx inference graph (graph JSON + dense model, embedding vectors as inputs) reproduces
    the evaluation predictions of every model family"""
    import os
    import hugectr_b200 as hugectr
    from hugectr_b200.models import build_deepfm, build_wdl
    from hugectr_b200.models.dlrm import build_dlrm_dcnv2
    from hugectr_b200.onnx.hugectr2onnx import convert
    B = 16
    kw = dict(comm=CPU(), max_eval_batches=1, batchsize_eval=B)
    m = {"deepfm": lambda: build_deepfm(batchsize=B, slot_sizes=[30] * 26, workspace_mb=1, **kw),
         "wdl": lambda: build_wdl(batchsize=B, wide_slot_sizes=[20, 30], deep_slot_sizes=[30] * 26,
                                  workspace_mb=(1, 1), **kw),
         "ncf": lambda: zoo.build_ncf("neumf", batchsize=B, num_users=100, num_items=80, **kw),
         "mmoe": lambda: zoo.build_mmoe(batchsize=B, num_slots=4, vocab=50, ev=8, expert_dims=(16, 8),
                                        tower_dim=8, **kw),
         "din": lambda: zoo.build_din(batchsize=B, seq_len=3, item_vocab=60, cate_vocab=10, user_vocab=20,
                                      ev=4, att_dims=(8, 4), mlp_dims=(8, 4), **kw),
         "bst": lambda: zoo.build_bst(batchsize=B, seq_len=3, item_vocab=60, user_vocab=20, ev=8, heads=2,
                                      ffn_dim=8, mlp_dims=(8, 4), **kw),
         "dcnv2": lambda: build_dlrm_dcnv2(batchsize=B, num_gpus=1, table_sizes=[50, 30, 20],
                                           multi_hot=[2, 1, 3], ev_size=8, mixed=False, bottom=(16, 8),
                                           top=(16, 1), cross_layers=1, projection_dim=4,
                                           use_cuda_graph=False, **kw)}[name]()
    for c in m.dense_layers:
        if c.layer_type == hugectr.Layer_t.Dropout:
            c.dropout_rate = 0.0
    m.compile()
    m.train()
    pre = str(tmp_path / name)
    m.save_params_to_files(pre, 1)
    m.graph_to_json(pre + ".json")
    m.eval()
    pred = [ll.pred.clone().reshape(-1) for ll in m.net_eval.loss_layers][-1]
    g = convert(pre + ".onnx", pre + ".json", pre + "_dense_1.model", convert_embedding=False, batch_size=B)
    ins = [torch.from_numpy(m.check_out_tensor(l["top"], hugectr.Tensor_t.Evaluate)) for l in g.emb_layers]
    for e in getattr(m, "ebcs_eval", []):
        ins += [e.top_data[tp["name"]].reshape(B, -1).float().clone() for tp in e.tops]
    dense = m.net_eval.tensors[m.input.dense_name].data.clone() if m.input.dense_dim > 0 else torch.zeros(B, 0)
    with torch.no_grad():
        out = g(dense, *ins).reshape(-1)
    torch.testing.assert_close(out, pred, atol=1e-6, rtol=1e-5)
    assert os.path.exists(pre + ".onnx") or os.path.exists(pre + ".onnx.pt")


def test_shared_bottom_multitask_uses_an_automatic_fanout():
    m = zoo.build_shared_bottom(batchsize=64, num_slots=6, vocab=100, ev=8, shared_dims=(32, 16), tower_dim=8,
                                comm=CPU(), max_eval_batches=1)
    losses = _train(m, 12)
    assert len(m.net_train.loss_layers) == 2 and len(losses) == 12
    assert any(l.cfg.layer_type.name == "Slice" for l in m.net_train.layers)     # inserted by the compiler


def test_criteo_dnn():
    losses = _train(zoo.build_criteo_dnn(batchsize=64, num_slots=5, vocab=200, ev=8, hidden=(32, 16),
                                         comm=CPU(), max_eval_batches=1), 12)
    assert len(losses) == 12          # random synthetic labels: finite losses are what _train checks


def test_dlrm_dcnv2_with_ftrl():
    import hugectr_b200 as hugectr
    m = zoo.build_dlrm_ftrl(batchsize=32, num_gpus=1, table_sizes=[50, 40, 30], multi_hot=[2, 1, 1], ev_size=8,
                            mixed=False, bottom=(16, 8), top=(16, 1), projection_dim=4, cross_layers=1,
                            comm=CPU(), max_eval_batches=1)
    assert m.opt_params.optimizer_type == hugectr.Optimizer_t.Ftrl
    losses = _train(m, 10)
    assert np.isfinite(losses).all()
