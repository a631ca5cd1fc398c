"""CPU numerics: every layer's fprop/bprop vs an independent PyTorch autograd oracle
(pattern of test/utest/core23_layer_test/*.cpp: random host data -> op -> reference)."""
import math

import pytest
import torch

import hugectr_b200 as hugectr
from hugectr_b200.layers import LAYER_REGISTRY, BuildCtx, ParamArena, TensorBag
from hugectr_b200.solver import CreateSolver, DenseLayer

L = hugectr.Layer_t


def build(layer_type, in_shapes, tops=("out",), dtype=torch.float32, int_inputs=(), **kw):
    arena = ParamArena()
    ctx = BuildCtx(arena, torch.device("cpu"), dtype, in_shapes[0][0], True, CreateSolver(), False)
    ins = []
    for i, s in enumerate(in_shapes):
        t = TensorBag(f"in{i}", s, dtype)
        t.data = torch.randn(s) * 0.5
        t.grad = torch.zeros(s)
        ins.append(t)
    cfg = DenseLayer(layer_type, [t.name for t in ins], list(tops), **kw)
    layer = LAYER_REGISTRY[layer_type](cfg, ins, ctx)
    arena.finalize(torch.device("cpu"), False)
    layer.allocate()
    arena.init_params(3)
    return layer, ins, arena


def check(layer, ins, ref_fn, atol=1e-4, params=None):
    layer.fprop(True)
    xs = [t.data.clone().requires_grad_(True) for t in ins]
    ws = [p.w.clone().requires_grad_(True) for p in layer.params]
    ys = ref_fn(*xs, *ws)
    if not isinstance(ys, (tuple, list)):
        ys = (ys,)
    for o, y in zip(layer.outputs, ys):
        torch.testing.assert_close(o.data, y.detach().reshape(o.data.shape), atol=atol, rtol=1e-3)
    gouts = [torch.randn_like(y) for y in ys]
    for o, g in zip(layer.outputs, gouts):
        o.grad.copy_(g.reshape(o.grad.shape))
    for p in layer.params:
        p.g.zero_()
    layer.bprop()
    grads = torch.autograd.grad(ys, xs + ws, gouts, allow_unused=True)
    for t, g in zip(ins, grads[:len(xs)]):
        if g is not None:
            torch.testing.assert_close(t.grad, g, atol=atol, rtol=1e-3)
    for p, g in zip(layer.params, grads[len(xs):]):
        if g is not None:
            torch.testing.assert_close(p.g, g.reshape(p.g.shape), atol=atol, rtol=1e-3)


def test_mlp():
    layer, ins, _ = build(L.MLP, [(16, 13)], num_outputs=[32, 16, 1],
                          activations=[hugectr.Activation_t.Relu, hugectr.Activation_t.Relu,
                                       hugectr.Activation_t.Non])

    def ref(x, w0, b0, w1, b1, w2, b2):
        h = torch.relu(x @ w0 + b0)
        h = torch.relu(h @ w1 + b1)
        return h @ w2 + b2
    check(layer, ins, ref)


def test_inner_product_3d():
    layer, ins, _ = build(L.InnerProduct, [(4, 5, 8)], num_output=6)
    check(layer, ins, lambda x, w, b: x @ w + b)


@pytest.mark.parametrize("proj", [0, 4])
def test_multicross(proj):
    layer, ins, _ = build(L.MultiCross, [(8, 12)], num_layers=2, projection_dim=proj)

    def ref_v2(x0, *ws):
        x = x0
        for l in range(2):
            U, V, b = ws[3 * l:3 * l + 3]
            x = x0 * ((x @ U) @ V + b) + x
        return x

    def ref_v1(x0, *ws):
        x = x0
        for l in range(2):
            w, b = ws[2 * l:2 * l + 2]
            x = x0 * (x @ w.reshape(-1, 1)) + b + x
        return x
    check(layer, ins, ref_v2 if proj else ref_v1)


def test_interaction():
    layer, ins, _ = build(L.Interaction, [(6, 8), (6, 5, 8)])

    def ref(mlp, emb):
        x = torch.cat([mlp.unsqueeze(1), emb], 1)
        z = torch.bmm(x, x.transpose(1, 2))
        li, lj = torch.tril_indices(6, 6, -1)
        return torch.cat([mlp, z[:, li, lj], torch.zeros(6, 1)], 1)
    check(layer, ins, ref)
    assert layer.outputs[0].shape == (6, 8 + 15 + 1)


def test_concat_slice_reshape():
    layer, ins, _ = build(L.Concat, [(4, 3), (4, 5)])
    check(layer, ins, lambda a, b: torch.cat([a, b], 1))
    layer, ins, _ = build(L.Slice, [(4, 10)], tops=("a", "b"), ranges=[(0, 6), (4, 10)])
    check(layer, ins, lambda x: (x[:, 0:6], x[:, 4:10]))
    layer, ins, _ = build(L.Reshape, [(4, 3, 5)], selected=True, selected_slots=[0, 2])
    check(layer, ins, lambda x: x[:, [0, 2], :].reshape(4, 10))


@pytest.mark.parametrize("lt,shape,kw,fn", [
    (L.ReLU, (4, 7), {}, torch.relu),
    (L.Sigmoid, (4, 7), {}, torch.sigmoid),
    (L.ELU, (4, 7), {"elu_alpha": 0.7}, lambda x: torch.nn.functional.elu(x, 0.7)),
    (L.FmOrder2, (4, 12), {"out_dim": 4},
     lambda x: 0.5 * (x.view(4, 3, 4).sum(1) ** 2 - (x.view(4, 3, 4) ** 2).sum(1))),
    (L.ReduceSum, (4, 7), {"axis": 1}, lambda x: x.sum(1, keepdim=True)),
    (L.ReduceMean, (4, 3, 5), {"axis": 1}, lambda x: x.mean(1, keepdim=True)),
    (L.Softmax, (4, 7), {}, lambda x: torch.softmax(x, -1)),
    (L.Scale, (4, 3), {"axis": 0, "factor": 2.0}, lambda x: x.repeat_interleave(2, 1)),
    (L.Scale, (4, 3), {"axis": 1, "factor": 3.0}, lambda x: x.repeat_interleave(3, 0)),
    (L.Select, (4, 6, 3), {"dim": 1, "index": [1, 4]}, lambda x: x[:, [1, 4]]),
    (L.Gather, (6, 5), {"indices": [0, 3, 5]}, lambda x: x[[0, 3, 5]]),
])
def test_unary(lt, shape, kw, fn):
    layer, ins, _ = build(lt, [shape], **kw)
    check(layer, ins, fn)


def test_binary_and_weighted():
    for lt, fn in [(L.Add, lambda a, b: a + b), (L.Sub, lambda a, b: a - b),
                   (L.ElementwiseMultiply, lambda a, b: a * b)]:
        layer, ins, _ = build(lt, [(4, 6), (4, 6)])
        check(layer, ins, fn)
    layer, ins, _ = build(L.WeightMultiply, [(4, 5)], weight_dims=[5, 3])
    check(layer, ins, lambda x, w: (x.unsqueeze(2) * w).reshape(4, 15))
    layer, ins, _ = build(L.MatrixMultiply, [(4, 3, 5), (4, 5, 2)])
    check(layer, ins, torch.matmul)
    layer, ins, _ = build(L.LayerNorm, [(4, 8)])
    check(layer, ins, lambda x, g, b: torch.nn.functional.layer_norm(x, (8,), g.reshape(-1), b.reshape(-1), 1e-5))


def test_mha_and_fused_reshape():
    layer, ins, _ = build(L.MultiHeadAttention, [(2, 4, 8), (2, 5, 8), (2, 5, 8)], num_attention_heads=2)

    def ref(q, k, v):
        qh, kh, vh = [t.view(2, -1, 2, 4).transpose(1, 2) for t in (q, k, v)]
        p = torch.softmax(qh @ kh.transpose(-1, -2) / 2.0, -1)
        return (p @ vh).transpose(1, 2).reshape(2, 4, 8)
    check(layer, ins, ref)
    layer, ins, _ = build(L.FusedReshapeConcat, [(3, 4, 2), (3, 4, 5)], tops=("his", "item"))
    check(layer, ins, lambda a, b: (torch.cat([a, b], 2)[:, :3].reshape(9, 7), torch.cat([a, b], 2)[:, 3]))


def test_bce_loss_semantics():
    arena = ParamArena()
    s = CreateSolver()
    ctx = BuildCtx(arena, torch.device("cpu"), torch.float32, 8, True, s, False)
    x = TensorBag("logit", (8, 1), torch.float32)
    x.data = torch.randn(8, 1)
    x.grad = torch.zeros(8, 1)
    y = TensorBag("label", (8, 1), torch.float32)
    y.data = (torch.rand(8, 1) > 0.5).float()
    y.needs_grad = False
    layer = LAYER_REGISTRY[L.BinaryCrossEntropyLoss](
        DenseLayer(L.BinaryCrossEntropyLoss, ["logit", "label"], ["loss"]), [x, y], ctx)
    layer.allocate()
    layer.fprop(True)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(x.data, y.data)
    torch.testing.assert_close(layer.outputs[0].data[0], ref, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(x.grad, (torch.sigmoid(x.data) - y.data) / 8, atol=1e-6, rtol=1e-5)
    layer.fprop(False)
    torch.testing.assert_close(layer.pred, torch.sigmoid(x.data))
