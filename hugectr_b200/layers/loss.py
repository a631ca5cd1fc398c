"""Loss layers + regularizers.

Semantics follow HugeCTR/src/loss.cu:158-403 (Appendix A.1 of SURVEY): in train mode the loss layer
emits the gradient of its logit input already divided by the *global* batch
(scaler / b_local / total_gpu_count), so the data-parallel all-reduce is a plain sum; in eval mode
it emits predictions (sigmoid / softmax) that the metrics read.  Regularizers
(HugeCTR/src/regularizers/*.cu) add  lambda*|w|_1 / 0.5*lambda*|w|_2^2  to the loss and initialise
wgrad with lambda*sign(w) / lambda*w before bprop.
"""
from __future__ import annotations

import torch

from ..enums import Regularizer_t
from ..ops import dense as D
from .base import Layer


class Regularizer:
    def __init__(self, kind, lambda_, params, batch):
        self.kind, self.lam, self.params, self.batch = kind, float(lambda_), params, batch

    def rterm(self) -> torch.Tensor:
        tot = None
        for p in self.params:
            # l1_regularizer.cu:47 (lambda / batch), l2_regularizer.cu:45 (lambda / (2 * batch))
            if self.kind == Regularizer_t.L1:
                v = p.w.abs().sum() * (self.lam / self.batch)
            else:
                v = (p.w * p.w).sum() * (self.lam * 0.5 / self.batch)
            tot = v if tot is None else tot + v
        return tot

    def init_wgrad(self):
        # reference: regularizer_initialize_wgrad (core23_network.cpp:59); scaled like the loss grad
        for p in self.params:
            if self.kind == Regularizer_t.L1:
                p.g.add_(torch.sign(p.w) * (self.lam / self.batch))
            else:
                p.g.add_(p.w * (self.lam / self.batch))


class LossLayer(Layer):
    is_loss = True

    def __init__(self, cfg, inputs, ctx):
        super().__init__(cfg, inputs, ctx)
        self.loss_weight = 1.0
        self.total_gpus = max(1, ctx.solver.num_gpus)
        self.scaler = float(ctx.solver.scaler) if ctx.mixed else 1.0
        self.gen_loss = bool(ctx.solver.gen_loss_summary)
        o = self._out(0, (1,), torch.float32)
        o.needs_grad = False
        self.regularizers = []

    def allocate(self):
        super().allocate()
        x = self.inputs[0]
        # eval predictions (what metrics read as "pred")
        self.pred = torch.zeros(x.shape, dtype=x.dtype, device=self.ctx.device)

    def _rterm(self):
        if not self.regularizers:
            return None
        tot = None
        for r in self.regularizers:
            v = r.rterm()
            tot = v if tot is None else tot + v
        return tot


class BinaryCrossEntropyLossLayer(LossLayer):
    """loss.cu:231-264.  bottom = [logits [b,1], label [b,1]]."""

    def __init__(self, cfg, inputs, ctx):
        super().__init__(cfg, inputs, ctx)
        assert inputs[0].shape[-1] == 1, "The feature dimension of BCE loss input should be 1"

    def fprop(self, is_train):
        x, y = self.inputs[0], self.inputs[1]
        n = x.data.numel()
        loss = self.outputs[0].data
        loss.zero_()
        if is_train and x.grad is not None:
            gs = self.scaler / n / self.total_gpus * self.loss_weight
            D.bce_loss(x.data, y.data, x.grad, loss, gs, self.loss_weight / n, True, self.gen_loss)
        else:
            D.bce_loss(x.data, y.data, self.pred, loss, 0.0, self.loss_weight / n, False,
                       self.gen_loss)
        r = self._rterm()
        if r is not None:
            loss.add_(r * self.loss_weight)

    def bprop(self):
        pass  # gradient already produced in fprop (fused)


class CrossEntropyLossLayer(LossLayer):
    """loss.cu:158-229: 2-class softmax CE over logits [b,2]; label<0.5 -> class 0."""

    def fprop(self, is_train):
        x, y = self.inputs[0], self.inputs[1]
        z = x.data.float()
        b = z.shape[0]
        a = torch.softmax(z[:, :2], dim=1)
        noclick = (y.data.reshape(-1).float() < 0.5)
        tgt = torch.stack([noclick.float(), (~noclick).float()], 1)
        loss = -(torch.log(torch.where(noclick, a[:, 0], a[:, 1]))).sum() / b
        r = self._rterm()
        if r is not None:
            loss = loss + r
        self.outputs[0].data.copy_((loss * self.loss_weight).reshape(1))
        if is_train and x.grad is not None:
            g = torch.zeros_like(z)
            g[:, :2] = (a - tgt) / b * self.scaler / self.total_gpus * self.loss_weight
            x.grad.copy_(g.to(x.grad.dtype))
        else:
            p = torch.zeros_like(z)
            p[:, :2] = a
            self.pred.copy_(p.to(self.pred.dtype))

    def bprop(self):
        pass


class MultiCrossEntropyLossLayer(LossLayer):
    """loss.cu:306-403: per-label weighted BCE; label < -0.5 marks a missing label."""

    def __init__(self, cfg, inputs, ctx):
        super().__init__(cfg, inputs, ctx)
        k = inputs[0].shape[-1]
        tw = list(cfg.target_weight_vec) or [1.0] * k
        assert len(tw) == k, "target_weight_vec must match the label dim"
        self.tw_list = tw

    def allocate(self):
        super().allocate()
        self.tw = torch.tensor(self.tw_list, dtype=torch.float32, device=self.ctx.device)

    def fprop(self, is_train):
        x, y = self.inputs[0], self.inputs[1]
        z = x.data.float()
        t = y.data.float().reshape(z.shape)
        size = z.numel()
        valid = (t >= -0.5).float()
        l = (torch.clamp(z, min=0) - z * t + torch.log1p(torch.exp(-z.abs()))) * self.tw * valid
        loss = l.sum() / size
        r = self._rterm()
        if r is not None:
            loss = loss + r
        self.outputs[0].data.copy_((loss * self.loss_weight).reshape(1))
        sig = torch.sigmoid(z)
        if is_train and x.grad is not None:
            g = (sig - t) * self.tw * valid / size * self.scaler / self.total_gpus * self.loss_weight
            x.grad.copy_(g.to(x.grad.dtype))
        else:
            self.pred.copy_(sig.to(self.pred.dtype))

    def bprop(self):
        pass
