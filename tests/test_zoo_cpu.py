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
