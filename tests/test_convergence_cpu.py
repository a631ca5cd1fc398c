"""Learning sanity: DLRM-DCNv2 must learn a synthetic teacher (labels are a hidden function of the keys
and dense features) -- checks that the whole forward / backward / optimizer loop carries signal."""
import torch

from hugectr_b200.data.batch import HostBatch
from hugectr_b200.metrics import auc_exact
from hugectr_b200.models.dlrm import build_dlrm_dcnv2
from hugectr_b200.parallel.comm import Comm


def teacher_batches(sizes, hot, B, seed=1, key_cap=None):
    gen = torch.Generator().manual_seed(seed)
    teach = [torch.randn(s, generator=gen) for s in sizes]
    wd = torch.randn(13, generator=gen)

    def batch():
        keys = [torch.randint(0, min(s, key_cap or s), (B, h), generator=gen) for s, h in zip(sizes, hot)]
        dense = torch.rand(B, 13, generator=gen)
        score = sum(t[k].sum(1) for t, k in zip(teach, keys)) + (dense - 0.5) @ wd
        return HostBatch((score > 0).float().view(B, 1), dense,
                         torch.cat([k.reshape(-1) for k in keys]).int(), None, B)
    return batch


def eval_auc(m, batch, n=8):
    ps, ys = [], []
    for _ in range(n):
        hb = batch()
        m._load_batch(hb, False)
        for e in m.ebcs_eval:
            e.forward(False)
        m.net_eval.fprop(False)
        ps.append(m.net_eval.loss_layers[0].pred.float().cpu().clone())
        ys.append(hb.label)
    return auc_exact(torch.cat(ps), torch.cat(ys))


def test_dcnv2_learns_a_teacher_on_cpu():
    sizes, hot, B = [200, 50, 400, 30], [2, 1, 3, 1], 256
    m = build_dlrm_dcnv2(batchsize=B, num_gpus=1, table_sizes=sizes, multi_hot=hot, ev_size=8, lr=0.05,
                         mixed=False, optimizer="adagrad", bottom=(16, 8), top=(32, 16, 1), cross_layers=2,
                         projection_dim=4, comm=Comm.single(torch.device("cpu")), use_cuda_graph=False,
                         batchsize_eval=B)
    m.compile()
    batch = teacher_batches(sizes, hot, B)
    auc0 = eval_auc(m, batch, 4)
    for _ in range(300):
        m.train_on_host_batch(batch())
    auc1 = eval_auc(m, batch)
    assert auc0 < 0.7 and auc1 > 0.97, (auc0, auc1)
