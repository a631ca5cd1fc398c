"""Per-GPU dense network: layer list built from DenseLayer configs over named tensors.

Reference: Network (HugeCTR/src/core23_network.cpp:33-404), graph analysis with automatic Slice
insertion for fan-out (HugeCTR/src/pybind/model_compile.cpp:624-684), top/bottom layer split for
embedding overlap (add_dense_layer_helpers.cpp:854-862).
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .enums import Layer_t
from .layers import LAYER_REGISTRY, LOSS_LAYERS, BuildCtx, Layer, Regularizer, TensorBag
from .solver import DenseLayer


def insert_fanout_slices(layer_cfgs: List[DenseLayer], source_names: List[str],
                         dims: Dict[str, int] = None) -> List[DenseLayer]:
    """A tensor consumed by more than one layer gets an explicit Slice fan-out (<= 5 branches in
    the reference; unbounded here).  Returns a new layer list with rewritten bottom names."""
    consumers: Dict[str, int] = {}
    for c in layer_cfgs:
        for b in c.bottom_names:
            consumers[b] = consumers.get(b, 0) + 1
    multi = {n for n, k in consumers.items() if k > 1}
    # labels feeding several losses are read-only: no fan-out needed
    out: List[DenseLayer] = []
    counters: Dict[str, int] = {}
    produced = set(source_names)
    pending_slice = {}

    def emit_slice(name):
        k = consumers[name]
        tops = [f"{name}_slice{i}" for i in range(k)]
        sl = DenseLayer(Layer_t.Slice, [name], tops, ranges=[(0, -1)] * k)
        sl._auto = True
        out.append(sl)
        pending_slice[name] = tops
        counters[name] = 0

    for n in source_names:
        if n in multi:
            emit_slice(n)
    for c in layer_cfgs:
        nb = []
        for b in c.bottom_names:
            if b in pending_slice:
                nb.append(pending_slice[b][counters[b]])
                counters[b] += 1
            else:
                nb.append(b)
        c2 = c
        if nb != c.bottom_names:
            import copy
            c2 = copy.copy(c)
            c2.bottom_names = nb
        out.append(c2)
        for t in c.top_names:
            produced.add(t)
            if t in multi:
                emit_slice(t)
    return out


class Network:
    def __init__(self, ctx: BuildCtx, sources: Dict[str, TensorBag], emb_tops=()):
        self.ctx = ctx
        self.tensors: Dict[str, TensorBag] = dict(sources)
        self.layers: List[Layer] = []
        self.loss_layers: List[Layer] = []
        self.emb_dependent = set(emb_tops)
        self.bottom_layers: List[Layer] = []   # independent of embeddings (overlappable)
        self.top_layers: List[Layer] = []

    def add_layer(self, cfg: DenseLayer) -> Layer:
        cls = LAYER_REGISTRY.get(cfg.layer_type)
        if cls is None:
            raise ValueError(f"unsupported layer type {cfg.layer_type}")
        ins = []
        for b in cfg.bottom_names:
            if b not in self.tensors:
                raise KeyError(f"bottom tensor '{b}' of layer {cfg.layer_type.name} is not defined")
            ins.append(self.tensors[b])
        if cfg.layer_type == Layer_t.Slice:
            w = ins[0].shape[-1]
            cfg.ranges = [(a, w if b_ == -1 else b_) for (a, b_) in cfg.ranges]
        self.ctx.layer_index = len(self.layers)
        layer = cls(cfg, ins, self.ctx)
        for o in layer.outputs:
            if o.name in self.tensors:
                raise ValueError(f"tensor name '{o.name}' defined twice")
            self.tensors[o.name] = o
        self.layers.append(layer)
        dep = any(b in self.emb_dependent for b in cfg.bottom_names)
        if dep:
            for t in cfg.top_names:
                self.emb_dependent.add(t)
            self.top_layers.append(layer)
        else:
            self.bottom_layers.append(layer)
        if cfg.layer_type in LOSS_LAYERS:
            self.loss_layers.append(layer)
        if cfg.use_regularizer and layer.params:
            layer.regularizer = (cfg.regularizer_type, cfg.lambda_)
        return layer

    def finalize(self):
        # gradient requirements: sources without grad storage never get one
        for layer in self.layers:
            layer.allocate()
            layer.post_allocate()
        # attach regularizers to the loss layers
        regs = []
        for layer in self.layers:
            if layer.regularizer is not None:
                kind, lam = layer.regularizer
                regs.append(Regularizer(kind, lam, layer.params, self.ctx.batch))
        for ll in self.loss_layers:
            ll.regularizers = regs
        self.regularizers = regs

    # ------------------------------------------------------------------ execution
    def fprop(self, is_train: bool, which: str = "all"):
        layers = self.layers if which == "all" else (self.bottom_layers if which == "bottom"
                                                     else self.top_layers)
        for layer in layers:
            layer.fprop(is_train)

    def bprop(self, which: str = "all"):
        layers = self.layers if which == "all" else (self.bottom_layers if which == "bottom"
                                                     else self.top_layers)
        if which in ("all", "top"):
            for r in getattr(self, "regularizers", []):
                r.init_wgrad()
        hook = getattr(self, "wgrad_hook", None)
        for layer in reversed(layers):
            layer.bprop()
            if hook is not None and layer.params:
                lo = min(p.offset for p in layer.params)
                hook(lo, self._range_end(max(p.offset for p in layer.params)))

    def _range_end(self, last_offset: int) -> int:
        """arena offset where the parameter after ``last_offset`` starts (padding belongs to the
        parameter in front of it), or the arena size"""
        import bisect
        offs = getattr(self, "_arena_offsets", None)
        arena = self.ctx.arena
        if offs is None:
            offs = self._arena_offsets = sorted(p.offset for p in arena.params)
        i = bisect.bisect_right(offs, last_offset)
        return offs[i] if i < len(offs) else int(arena.wgrad.numel())

    def loss_value(self) -> torch.Tensor:
        tot = None
        for ll in self.loss_layers:
            v = ll.outputs[0].data
            tot = v.clone() if tot is None else tot + v
        return tot

    def summary_rows(self):
        rows = []
        for layer in self.layers:
            cfg = layer.cfg
            rows.append((cfg.layer_type.name, ",".join(cfg.bottom_names), ",".join(cfg.top_names),
                         ";".join(str(tuple(o.shape)) for o in layer.outputs)))
        return rows
