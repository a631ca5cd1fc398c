"""GPU learning sanity (runs last): the smoke-sized DLRM-DCNv2 in bf16 with the fused kernels and a CUDA
graph must learn the synthetic teacher like the fp32 CPU model does (CPU reference: AUC 0.99 after 150
steps with this configuration)."""
import pytest
import torch

from test_convergence_cpu import eval_auc, teacher_batches

pytestmark = pytest.mark.gpu


def test_dcnv2_learns_a_teacher_on_gpu_bf16_graph():
    from hugectr_b200.models.dlrm import build_dlrm_dcnv2
    sizes, hot, B = [5000, 300, 20000, 40, 1000, 7], [3, 1, 10, 2, 1, 1], 256
    m = build_dlrm_dcnv2(batchsize=B, num_gpus=1, table_sizes=sizes, multi_hot=hot, ev_size=128, lr=0.005,
                         mixed=True, bottom=(512, 256, 128), top=(1024, 512, 256, 1), cross_layers=3,
                         projection_dim=512, use_cuda_graph=True, batchsize_eval=B)
    m.compile()
    batch = teacher_batches(sizes, hot, B, key_cap=200)
    for _ in range(150):
        m.train_on_host_batch(batch())
    torch.cuda.synchronize()
    auc = eval_auc(m, batch, 4)
    assert auc > 0.9, auc
