"""Model API surface on CPU: freeze / unfreeze, check_out_tensor, learning-rate control, training
callbacks and early stop, epoch mode, multi-loss weights (model_wrapper.hpp:133-221)."""
import numpy as np
import pytest
import torch

import hugectr_b200 as hugectr
from hugectr_b200.models import build_dcn, zoo
from hugectr_b200.parallel.comm import Comm

CPU = lambda: Comm.single(torch.device("cpu"))


def _dcn(**kw):
    m = build_dcn(batchsize=64, slot_sizes=[60] * 26, workspace_mb=1, comm=CPU(), max_eval_batches=1, **kw)
    for c in m.dense_layers:
        if c.layer_type == hugectr.Layer_t.Dropout:
            c.dropout_rate = 0.0
    return m


def test_freeze_and_unfreeze():
    m = _dcn()
    m.compile()
    m.train()
    rt = m.legacy_train[0]
    w0, t0 = m.arena.weights.clone(), rt.table.clone()
    m.freeze_dense()
    m.train()
    assert torch.equal(m.arena.weights, w0) and not torch.equal(rt.table, t0)
    t1 = rt.table.clone()
    m.unfreeze_dense()
    m.freeze_embedding()
    m.train()
    assert not torch.equal(m.arena.weights, w0) and torch.equal(rt.table, t1)
    m.unfreeze_embedding()
    m.train()
    assert not torch.equal(rt.table, t1)


def test_check_out_tensor_and_learning_rate_control():
    m = _dcn(lr=0.01, warmup_steps=4)
    m.compile()
    m.train()
    x = m.check_out_tensor("concat1", hugectr.Tensor_t.Train)
    assert isinstance(x, np.ndarray) and x.shape == (64, 26 * 16 + 13)
    with pytest.raises(KeyError):
        m.check_out_tensor("nope")
    assert abs(float(m.lr_t) - 0.01 / 4) < 1e-9                  # step 1 of a 4-step warm-up
    m.set_learning_rate(0.5)
    m.reset_learning_rate_scheduler(0.5, warmup_steps=1)
    m.train()
    assert abs(float(m.lr_t) - 0.5) < 1e-9
    sch = m.get_learning_rate_scheduler()
    assert sch.get_next() > 0


def test_callbacks_can_stop_training_and_epoch_mode():
    events = []

    class CB(hugectr.TrainingCallback):
        def on_training_start(self):
            events.append("start")

        def on_eval_end(self, it, res):
            events.append(("eval", it, sorted(res)))
            return it >= 4                                  # stop after the second evaluation

        def on_training_end(self, it):
            events.append(("end", it))

    m = _dcn(training_callbacks=[CB()])
    m.compile()
    done = m.fit(max_iter=50, display=100, eval_interval=2, snapshot=0)
    assert events[0] == "start" and events[-1] == ("end", done) and done == 4
    assert [e[1] for e in events if e[0] == "eval"] == [2, 4] and events[1][2] == ["AUC"]
    # epoch mode over a finite synthetic source: every epoch ends when the reader runs dry
    m2 = _dcn()
    m2.compile()
    m2.reader_train.num_batches = 3
    it = m2.fit(num_epochs=2, display=100, eval_interval=0, snapshot=0)
    assert it == 6


def test_compile_with_loss_weights():
    m = zoo.build_mmoe(batchsize=32, num_slots=4, vocab=50, ev=8, expert_dims=(16, 8), tower_dim=8,
                       comm=CPU(), max_eval_batches=1)
    m.compile(loss_names=["loss0", "loss1"], loss_weights=[0.25, 2.0])
    assert [l.loss_weight for l in m.net_train.loss_layers] == [0.25, 2.0]
    m.train()
    total = m.get_current_loss()
    parts = [float(l.outputs[0].data) for l in m.net_train.loss_layers]
    assert abs(total - sum(parts)) < 1e-5


def test_get_eval_metrics_after_fit_returns_the_last_evaluation():
    """`fit` finalises the metrics at every evaluation; asking again afterwards returns those values rather than the
    result of finalising empty accumulators; new `eval()` batches start a new evaluation"""
    import torch
    from hugectr_b200.models.dlrm import build_dlrm_dcnv2
    from hugectr_b200.parallel.comm import Comm
    m = build_dlrm_dcnv2(batchsize=32, batchsize_eval=48, num_gpus=1, table_sizes=[50, 20], multi_hot=[2, 1], ev_size=8,
                         mixed=False, bottom=(16, 8), top=(16, 1), projection_dim=4, cross_layers=1, use_cuda_graph=False,
                         comm=Comm.single(torch.device("cpu")), max_eval_batches=3)
    m.compile()
    m.fit(max_iter=4, display=2, eval_interval=2, snapshot=10**9)
    a = dict(m.get_eval_metrics())
    assert 0.0 < a["AUC"] < 1.0 and dict(m.get_eval_metrics()) == a          # stable until new batches arrive
    for _ in range(2):
        assert m.eval()
    b = dict(m.get_eval_metrics())
    assert 0.0 < b["AUC"] < 1.0 and b != a
