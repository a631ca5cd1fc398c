"""CPU numerics: every layer's fprop/bprop vs an independent PyTorch autograd oracle
(pattern of test/utest/core23_layer_test/*.cpp: random host data -> op -> reference)."""

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


# ---------------------------------------------------------------------------- remaining layer types
def test_batchnorm_train_and_eval():
    layer, ins, _ = build(L.BatchNorm, [(16, 6)], factor=0.9, eps=1e-5)
    bn = torch.nn.BatchNorm1d(6, eps=1e-5, momentum=0.1)
    bn.train()

    def ref(x, g, b):
        return torch.nn.functional.batch_norm(x, None, None, g.reshape(-1), b.reshape(-1), True, 0.1, 1e-5)
    check(layer, ins, ref)
    # running statistics (momentum = factor) then inference with them
    x = ins[0].data
    m = x.mean(0) * 0.1
    v = 0.9 + 0.1 * x.var(0, unbiased=False)
    torch.testing.assert_close(layer.state["mean"], m, atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(layer.state["var"], v, atol=1e-5, rtol=1e-4)
    layer.fprop(False)
    g, b = layer.params[0].w.reshape(-1), layer.params[1].w.reshape(-1)
    torch.testing.assert_close(layer.outputs[0].data, (x - m) / torch.sqrt(v + 1e-5) * g + b,
                               atol=1e-4, rtol=1e-3)


def test_prelu_dice_and_dropout_and_cast():
    layer, ins, _ = build(L.PReLU_Dice, [(32, 5)], elu_alpha=0.2, eps=1e-8)

    def dice(x):
        ex = x.mean(0, keepdim=True)
        var = (x * x).mean(0, keepdim=True) - ex * ex
        p = torch.sigmoid((x - ex) / torch.sqrt(var + 1e-8))
        return p * x + (1 - p) * 0.2 * x
    check(layer, ins, dice)
    # dropout: inverted scaling, mask reused by bprop, identity in eval
    layer, ins, _ = build(L.Dropout, [(64, 50)], dropout_rate=0.25)
    layer.fprop(True)
    y, x = layer.outputs[0].data, ins[0].data
    kept = y != 0
    torch.testing.assert_close(y[kept], x[kept] / 0.75)
    assert 0.6 < kept.float().mean().item() < 0.9
    layer.outputs[0].grad.fill_(1.0)
    layer.bprop()
    torch.testing.assert_close(ins[0].grad, kept.float() / 0.75)
    layer.fprop(False)
    torch.testing.assert_close(layer.outputs[0].data, x)
    # cast: value preserving round trip through the compute dtype
    layer, ins, _ = build(L.Cast, [(4, 6)])
    layer.fprop(True)
    torch.testing.assert_close(layer.outputs[0].data.float(), ins[0].data, atol=1e-2, rtol=1e-2)


def test_concat3d_dotproduct_reluhalf_general_reshape_concat():
    layer, ins, _ = build(L.Concat3D, [(4, 2, 5), (4, 3, 5)], axis=1)
    check(layer, ins, lambda a, b: torch.cat([a, b], 1))
    layer, ins, _ = build(L.Concat3D, [(4, 3, 2), (4, 3, 6)], axis=2)
    check(layer, ins, lambda a, b: torch.cat([a, b], 2))
    layer, ins, _ = build(L.DotProduct, [(4, 7), (4, 7)])
    check(layer, ins, lambda a, b: a * b)
    layer, ins, _ = build(L.ReLUHalf, [(4, 7)])
    check(layer, ins, torch.relu)
    layer, ins, _ = build(L.FusedReshapeConcatGeneral, [(3, 4, 2), (3, 4, 5)], tops=("o",))
    check(layer, ins, lambda a, b: torch.cat([a, b], 2).reshape(12, 7))


def test_masked_softmax_and_sequence_mask():
    arena = ParamArena()
    ctx = BuildCtx(arena, torch.device("cpu"), torch.float32, 4, True, CreateSolver(), False)
    lf = TensorBag("lf", (4, 1), torch.float32)
    lt = TensorBag("lt", (4, 1), torch.float32)
    lf.data = torch.tensor([[1.], [3.], [2.], [0.]])
    lt.data = torch.tensor([[2.], [1.], [3.], [3.]])
    lf.needs_grad = lt.needs_grad = False
    sm = LAYER_REGISTRY[L.SequenceMask](DenseLayer(L.SequenceMask, ["lf", "lt"], ["mask"],
                                                   max_sequence_len_from=3, max_sequence_len_to=3),
                                        [lf, lt], ctx)
    sm.allocate()
    sm.fprop(True)
    m = sm.outputs[0].data
    assert m.shape == (4, 1, 3, 3)
    exp1 = torch.zeros(3, 3)
    exp1[:3, :1] = 1            # sample 1: from-length 3, to-length 1
    torch.testing.assert_close(m[1, 0], exp1)
    assert m[3].sum() == 0      # from-length 0
    # masked softmax: masked positions get (numerically) zero probability, rows still sum to one
    x = TensorBag("x", (4, 1, 3, 3), torch.float32)
    x.data = torch.randn(4, 1, 3, 3)
    x.grad = torch.zeros(4, 1, 3, 3)
    ms = LAYER_REGISTRY[L.MaskedSoftmax](DenseLayer(L.MaskedSoftmax, ["x", "mask"], ["p"]),
                                         [x, sm.outputs[0]], ctx)
    ms.allocate()
    ms.fprop(True)
    p = ms.outputs[0].data
    ref = torch.softmax(torch.where(m > 0, x.data, torch.full_like(x.data, -10000.0)), -1)
    torch.testing.assert_close(p, ref, atol=1e-6, rtol=1e-5)
    assert (p[0, 0, 0, 2:] < 1e-6).all()


def test_gru_matches_torch():
    b, S, v, h = 3, 4, 5, 6
    layer, ins, _ = build(L.GRU, [(1, b * S * v)], batchsize=b, SeqLength=S, vector_size=v, num_output=h)
    layer.fprop(True)
    gru = torch.nn.GRU(v, h, batch_first=True)
    with torch.no_grad():
        gru.weight_ih_l0.copy_(layer.params[0].w)
        gru.weight_hh_l0.copy_(layer.params[1].w)
        gru.bias_ih_l0.copy_(layer.params[2].w.reshape(-1))
        gru.bias_hh_l0.copy_(layer.params[3].w.reshape(-1))
    x = ins[0].data.reshape(b, S, v).clone().requires_grad_(True)
    y, _ = gru(x)
    torch.testing.assert_close(layer.outputs[0].data.reshape(b, S, h), y.detach(), atol=1e-5, rtol=1e-4)
    g = torch.randn_like(y)
    layer.outputs[0].grad.copy_(g.reshape(1, -1))
    for p in layer.params:
        p.g.zero_()
    layer.bprop()
    y.backward(g)
    torch.testing.assert_close(ins[0].grad.reshape(b, S, v), x.grad, atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(layer.params[0].g, gru.weight_ih_l0.grad, atol=1e-5, rtol=1e-4)


def _loss_layer(lt, logits, labels, **kw):
    arena = ParamArena()
    ctx = BuildCtx(arena, torch.device("cpu"), torch.float32, logits.shape[0], True, CreateSolver(), False)
    x = TensorBag("logit", tuple(logits.shape), torch.float32)
    x.data = logits.clone()
    x.grad = torch.zeros_like(logits)
    y = TensorBag("label", tuple(labels.shape), torch.float32)
    y.data = labels.clone()
    y.needs_grad = False
    layer = LAYER_REGISTRY[lt](DenseLayer(lt, ["logit", "label"], ["loss"], **kw), [x, y], ctx)
    layer.allocate()
    layer.fprop(True)
    return layer, x


def test_cross_entropy_and_multi_cross_entropy_losses():
    torch.manual_seed(1)
    z = torch.randn(8, 2)
    lab = (torch.rand(8, 1) > 0.5).float()
    layer, x = _loss_layer(L.CrossEntropyLoss, z, lab)
    zz = z.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(zz, lab.reshape(-1).long())
    ref.backward()
    torch.testing.assert_close(layer.outputs[0].data[0], ref.detach(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(x.grad, zz.grad, atol=1e-6, rtol=1e-5)
    # multi-label weighted BCE with a missing label (-1)
    z = torch.randn(6, 3)
    lab = (torch.rand(6, 3) > 0.5).float()
    lab[2, 1] = -1.0
    tw = [0.2, 0.5, 0.3]
    layer, x = _loss_layer(L.MultiCrossEntropyLoss, z, lab, target_weight_vec=tw)
    zz = z.clone().requires_grad_(True)
    valid = (lab > -0.5).float()
    per = torch.nn.functional.binary_cross_entropy_with_logits(zz, lab.clamp(min=0), reduction="none")
    ref = (per * valid * torch.tensor(tw)).sum() / (6 * 3)     # loss.cu:313,327: / (batch * labels)
    ref.backward()
    torch.testing.assert_close(layer.outputs[0].data[0], ref.detach(), atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(x.grad, zz.grad, atol=1e-6, rtol=1e-4)
