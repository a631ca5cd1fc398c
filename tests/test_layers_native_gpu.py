"""Native long-tail layer kernels (csrc/layers.cu) vs the fp32 PyTorch formulation of the same layer (the CPU
path): forward outputs, input gradients and parameter gradients, fp32 and bf16 storage."""
import pytest
import torch

import hugectr_b200 as hugectr
from hugectr_b200.layers import LAYER_REGISTRY, BuildCtx, ParamArena, TensorBag
from hugectr_b200.solver import CreateSolver, DenseLayer

L = hugectr.Layer_t
pytestmark = pytest.mark.gpu


def build(layer_type, in_shapes, dev, dtype, tops=("out",), masks=(), is_train=True, **kw):
    arena = ParamArena()
    ctx = BuildCtx(arena, dev, dtype, in_shapes[0][0], is_train, CreateSolver(), dtype == torch.bfloat16)
    g = torch.Generator().manual_seed(7)
    ins = []
    for i, s in enumerate(in_shapes):
        t = TensorBag(f"in{i}", s, dtype)
        if i in masks:
            t.data = (torch.rand(s, generator=g) > 0.3).to(dtype).to(dev)
            t.needs_grad = False
        else:
            t.data = (torch.randn(s, generator=g) * 0.5).to(dtype).to(dev)
            t.grad = torch.zeros(s, dtype=dtype, device=dev)
        ins.append(t)
    cfg = DenseLayer(layer_type, [t.name for t in ins], list(tops), **kw)
    layer = LAYER_REGISTRY[layer_type](cfg, ins, ctx)
    arena.finalize(dev, dtype == torch.bfloat16)
    layer.allocate()
    arena.init_params(3)
    return layer, ins, arena


CASES = [
    ("elu", L.ELU, [(64, 37)], {"elu_alpha": 0.7}),
    ("dice", L.PReLU_Dice, [(256, 24)], {"elu_alpha": 0.2, "eps": 1e-5}),
    ("fm", L.FmOrder2, [(128, 48)], {"out_dim": 16}),
    ("rsum", L.ReduceSum, [(64, 37)], {"axis": 1}),
    ("rmean", L.ReduceMean, [(32, 5, 24)], {"axis": 1}),
    ("softmax", L.Softmax, [(64, 50)], {}),
    ("scale0", L.Scale, [(16, 6)], {"axis": 0, "factor": 3.0}),
    ("scale1", L.Scale, [(16, 6)], {"axis": 1, "factor": 2.0}),
    ("select", L.Select, [(8, 6, 5)], {"dim": 1, "index": [1, 4, 4]}),
    ("gather", L.Gather, [(9, 7)], {"indices": [0, 3, 8, 3]}),
    ("add", L.Add, [(32, 20), (32, 20), (32, 20)], {}),
    ("sub", L.Sub, [(32, 20), (32, 20)], {}),
    ("mul", L.ElementwiseMultiply, [(32, 20), (32, 20), (32, 20)], {}),
    ("wmul", L.WeightMultiply, [(200, 7)], {"weight_dims": [7, 12]}),
    ("matmul3", L.MatrixMultiply, [(6, 20, 9), (6, 9, 33)], {}),
    ("matmul2", L.MatrixMultiply, [(40, 19), (19, 50)], {}),
    ("lnorm", L.LayerNorm, [(100, 48)], {}),
    ("bnorm", L.BatchNorm, [(300, 20)], {"factor": 0.9, "eps": 1e-5}),
    ("mha", L.MultiHeadAttention, [(4, 10, 32), (4, 12, 32), (4, 12, 32)], {"num_attention_heads": 4}),
    ("frc", L.FusedReshapeConcat, [(5, 6, 4), (5, 6, 8)], {"_tops": ("his", "item")}),
    ("frcg", L.FusedReshapeConcatGeneral, [(5, 6, 4), (5, 6, 8)], {}),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("name,lt,shapes,kw", CASES, ids=[c[0] for c in CASES])
def test_native_layer_matches_torch(name, lt, shapes, kw, dtype):
    kw = dict(kw)
    tops = kw.pop("_tops", ("out",))
    gpu, gins, garena = build(lt, shapes, torch.device("cuda"), dtype, tops=tops, **kw)
    cpu, cins, carena = build(lt, shapes, torch.device("cpu"), torch.float32, tops=tops, **kw)
    for a, b in zip(gins, cins):
        b.data = a.data.float().cpu()
    carena.weights.copy_(garena.weights.cpu())
    gpu.fprop(True)
    cpu.fprop(True)
    assert gpu._ran_native, "native kernels were not used on the GPU"
    tol = dict(atol=2e-5, rtol=1e-4) if dtype == torch.float32 else dict(atol=4e-2, rtol=4e-2)
    g = torch.Generator().manual_seed(11)
    for og, oc in zip(gpu.outputs, cpu.outputs):
        torch.testing.assert_close(og.data.float().cpu(), oc.data, **tol)
        dy = torch.randn(oc.data.shape, generator=g) * 0.5
        oc.grad.copy_(dy)
        og.grad.copy_(dy.to(dtype).cuda())
    for p in list(gpu.params) + list(cpu.params):
        p.g.zero_()
    gpu.bprop()
    cpu.bprop()
    for a, b in zip(gins, cins):
        if b.grad is not None:
            torch.testing.assert_close(a.grad.float().cpu(), b.grad, **tol)
    ptol = tol if dtype == torch.float32 else dict(atol=0.25, rtol=5e-2)
    for pg, pc in zip(gpu.params, cpu.params):
        torch.testing.assert_close(pg.g.cpu(), pc.g, **ptol)


def test_native_masked_softmax():
    for dtype in (torch.float32, torch.bfloat16):
        gpu, gins, _ = build(L.MaskedSoftmax, [(6, 2, 5, 9), (6, 2, 5, 9)], torch.device("cuda"), dtype, masks=(1,))
        cpu, cins, _ = build(L.MaskedSoftmax, [(6, 2, 5, 9), (6, 2, 5, 9)], torch.device("cpu"), torch.float32,
                             masks=(1,))
        for a, b in zip(gins, cins):
            b.data = a.data.float().cpu()
        gpu.fprop(True); cpu.fprop(True)
        assert gpu._ran_native
        tol = dict(atol=1e-5, rtol=1e-4) if dtype == torch.float32 else dict(atol=2e-2, rtol=2e-2)
        torch.testing.assert_close(gpu.outputs[0].data.float().cpu(), cpu.outputs[0].data, **tol)
        dy = torch.randn(6, 2, 5, 9)
        cpu.outputs[0].grad.copy_(dy); gpu.outputs[0].grad.copy_(dy.to(dtype).cuda())
        gpu.bprop(); cpu.bprop()
        torch.testing.assert_close(gins[0].grad.float().cpu(), cins[0].grad, **tol)


def test_native_batchnorm_eval_uses_running_stats():
    gpu, gins, _ = build(L.BatchNorm, [(64, 10)], torch.device("cuda"), torch.float32, factor=0.9, eps=1e-5)
    for _ in range(3):
        gpu.fprop(True)
    x = gins[0].data
    gpu.fprop(False)
    m, v = gpu.state["mean"], gpu.state["var"]
    exp = (x - m) / torch.sqrt(v + 1e-5) * gpu.params[0].w.reshape(-1) + gpu.params[1].w.reshape(-1)
    torch.testing.assert_close(gpu.outputs[0].data, exp, atol=1e-5, rtol=1e-4)
    bm = x.mean(0)
    assert (m - (1 - 0.9 ** 3) * bm).abs().max() < 1e-5     # momentum update of a constant batch mean


def test_native_gru_matches_torch():
    b, S, v, h = 5, 6, 7, 8
    layer, ins, _ = build(L.GRU, [(1, b * S * v)], torch.device("cuda"), torch.float32, batchsize=b, SeqLength=S,
                          vector_size=v, num_output=h)
    layer.fprop(True)
    assert layer._ran_native
    gru = torch.nn.GRU(v, h, batch_first=True).cuda()
    with torch.no_grad():
        gru.weight_ih_l0.copy_(layer.params[0].w)
        gru.weight_hh_l0.copy_(layer.params[1].w)
        gru.bias_ih_l0.copy_(layer.params[2].w.reshape(-1))
        gru.bias_hh_l0.copy_(layer.params[3].w.reshape(-1))
    x = ins[0].data.reshape(b, S, v).clone().requires_grad_(True)
    y, _ = gru(x)
    # (tolerances: fast-math exp / tanh in the gate kernel, TF32 inside cuDNN on the torch side)
    torch.testing.assert_close(layer.outputs[0].data.reshape(b, S, h), y.detach(), atol=2e-3, rtol=2e-2)
    g = torch.randn_like(y)
    layer.outputs[0].grad.copy_(g.reshape(1, -1))
    for p in layer.params:
        p.g.zero_()
    layer.bprop()
    y.backward(g)
    tol = dict(atol=5e-3, rtol=2e-2)
    torch.testing.assert_close(ins[0].grad.reshape(b, S, v), x.grad, **tol)
    torch.testing.assert_close(layer.params[0].g, gru.weight_ih_l0.grad, **tol)
    torch.testing.assert_close(layer.params[1].g, gru.weight_hh_l0.grad, **tol)
    torch.testing.assert_close(layer.params[2].g.reshape(-1), gru.bias_ih_l0.grad, **tol)
    torch.testing.assert_close(layer.params[3].g.reshape(-1), gru.bias_hh_l0.grad, **tol)
