"""Sparse Operation Kit, PyTorch flavour.

API parity with ``sparse_operation_kit`` (sparse_operation_kit/sparse_operation_kit/*.py):
``init``, ``Variable`` / ``DistributedVariable`` (row-wise key % N sharding), ``LocalizedVariable``
(whole table on one GPU), ``DynamicVariable`` (hash-backed, unbounded vocabulary; the role of
HKV / DET), ``lookup_sparse`` (lookup.py:425-541), ``all2all_dense_embedding``,
``OptimizerWrapper`` / ``SGD``, ``dump`` / ``load`` / ``incremental_model_dump`` (dump_load.py),
``filter_variables``.  The reference is a TensorFlow plugin that communicates through Horovod
(allgather of keys, alltoall of vectors); here variables are sharded torch tensors, communication is
``torch.distributed`` and the backward is a custom autograd Function that hands *sparse* gradients
(unique rows + reduced grads) to the optimizer wrapper.
"""
from __future__ import annotations

import os
import struct
import time
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from ..embedding.hashtable import HashTable
from ..embedding.ops import sparse_opt_reference
from ..enums import Optimizer_t
from ..parallel.comm import Comm

_comm: Optional[Comm] = None
_variables: List["Variable"] = []


def init(comm: Optional[Comm] = None, device: Optional[str] = None):
    """sok.init(): bind to the process group (one process per GPU)."""
    global _comm
    _comm = comm or Comm.init_from_env(device)
    return _comm


def _c() -> Comm:
    if _comm is None:
        init()
    return _comm


def rank(): return _c().rank
def num_gpus(): return _c().world_size


class Variable:
    """Static-shape embedding variable. mode='distributed': row r lives on rank r % N at local row
    r // N; mode='localized:<gpu>': the whole table lives on one GPU."""

    def __init__(self, initial_value=None, shape=None, dtype=torch.float32, mode: str = "distributed",
                 name: Optional[str] = None, initializer: str = "uniform", trainable: bool = True):
        c = _c()
        self.name = name or f"sok_var_{len(_variables)}"
        self.mode = mode
        self.trainable = trainable
        if initial_value is not None:
            full = torch.as_tensor(initial_value, dtype=dtype)
            shape = tuple(full.shape)
        else:
            full = None
        self.shape = tuple(shape)
        self.vocab, self.dim = self.shape
        self.world, self.rank = c.world_size, c.rank
        self.device = c.device
        if mode.startswith("localized"):
            self.owner = int(mode.split(":")[1]) if ":" in mode else 0
            rows = self.vocab if self.rank == self.owner else 0
            keys = torch.arange(self.vocab)
        else:
            self.owner = -1
            keys = torch.arange(self.rank, self.vocab, self.world)
            rows = keys.numel()
        if full is not None:
            local = full[keys] if rows else torch.zeros(0, self.dim)
        else:
            import zlib
            g = torch.Generator().manual_seed(zlib.crc32(self.name.encode()) & 0xFFFF)
            allv = (torch.rand(self.vocab, self.dim, generator=g) * 2 - 1) * 0.05
            local = allv[keys] if rows else torch.zeros(0, self.dim)
        self.weight = local.to(self.device, dtype).contiguous()
        self.states = {}
        self.sparse_grad = None
        self.last_touch = torch.zeros(self.weight.shape[0], dtype=torch.float64)
        _variables.append(self)

    # key -> (owner rank, local row)
    def locate(self, keys: torch.Tensor):
        if self.owner >= 0:
            return torch.full_like(keys, self.owner), keys
        return keys % self.world, torch.div(keys, self.world, rounding_mode="floor")

    def local_rows(self, keys: torch.Tensor, create: bool = True) -> torch.Tensor:
        _, r = self.locate(keys)
        return r

    def global_keys(self) -> torch.Tensor:
        n = self.weight.shape[0]
        if self.owner >= 0:
            return torch.arange(n)
        return torch.arange(n) * self.world + self.rank


DistributedVariable = Variable


def LocalizedVariable(*a, gpu: int = 0, **kw):
    kw["mode"] = f"localized:{gpu}"
    return Variable(*a, **kw)


class DynamicVariable(Variable):
    """Hash-backed variable with unbounded key space (sok.DynamicVariable over HKV/DET):
    rows are created on first lookup; capacity grows geometrically."""

    def __init__(self, dimension: int, var_type: str = "hbm", initializer: Union[str, float] = "uniform",
                 init_capacity: int = 1 << 16, max_capacity: int = 1 << 26, name: Optional[str] = None,
                 key_type=torch.int64, dtype=torch.float32, trainable: bool = True):
        c = _c()
        self.name = name or f"sok_dynvar_{len(_variables)}"
        self.mode = "dynamic"
        self.owner = -1
        self.trainable = trainable
        self.dim = int(dimension)
        self.world, self.rank, self.device = c.world_size, c.rank, c.device
        self.var_type = var_type
        self.initializer = initializer
        self.max_capacity = max_capacity
        self.hash = HashTable(max_capacity, self.device)
        self.weight = torch.zeros(init_capacity, self.dim, dtype=dtype, device=self.device)
        self._init_rows(0, init_capacity)
        self.states = {}
        self.sparse_grad = None
        self.last_touch = torch.zeros(init_capacity, dtype=torch.float64)
        self.vocab = -1
        self.shape = (-1, self.dim)
        _variables.append(self)

    def _init_rows(self, lo, hi):
        if isinstance(self.initializer, (int, float)):
            self.weight[lo:hi] = float(self.initializer)
        else:
            self.weight[lo:hi].uniform_(-0.05, 0.05)

    def _grow(self, need: int):
        cap = self.weight.shape[0]
        if need <= cap:
            return
        new = cap
        while new < need:
            new *= 2
        w = torch.zeros(new, self.dim, dtype=self.weight.dtype, device=self.device)
        w[:cap] = self.weight
        old = self.weight
        self.weight = w
        self._init_rows(cap, new)
        for k, s in list(self.states.items()):
            ns = torch.zeros(new, self.dim, dtype=s.dtype, device=self.device)
            ns[:cap] = s
            self.states[k] = ns
        lt = torch.zeros(new, dtype=torch.float64)
        lt[:cap] = self.last_touch
        self.last_touch = lt
        del old

    def local_rows(self, keys: torch.Tensor, create: bool = True) -> torch.Tensor:
        k = keys.to(self.device).to(torch.int64)
        rows = self.hash.get_insert(k) if create else self.hash.get(k)
        self._grow(self.hash.size())
        return rows

    def locate(self, keys):
        return keys % self.world, keys

    def global_keys(self):
        k, r = self.hash.dump()
        return k

    @property
    def size(self):
        return self.hash.size()


# ----------------------------------------------------------------------------- lookup
def _to_csr(sp_ids):
    """accept [b, H] padded (-1), or (values, row_lengths)"""
    if isinstance(sp_ids, (tuple, list)) and len(sp_ids) == 2 and sp_ids[0].dim() == 1:
        vals, lens = sp_ids
        return vals.to(torch.int64), lens.to(torch.int64)
    t = sp_ids.to(torch.int64)
    mask = t >= 0
    return t[mask], mask.sum(1)


class _Lookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, var, vals, lens, combiner, wts=None):
        c = _c()
        dev = var.device
        vals = vals.to(dev)
        lens = lens.to(dev)
        b = lens.numel()
        seg = torch.repeat_interleave(torch.arange(b, device=dev), lens)
        wts = None if wts is None else wts.to(dev).float().reshape(-1)
        ctx.w_own = None
        if c.world_size == 1:
            rows = var.local_rows(vals)
            vec = var.weight[rows].float()
            if wts is not None:
                vec = vec * wts.unsqueeze(1)
                ctx.w_own = wts
            out = torch.zeros(b, var.dim, device=dev).index_add_(0, seg, vec)
            ctx.rows, ctx.seg_owner = rows, None
        else:
            # allgather keys (+segment ids, + source rank), local lookup of owned keys, alltoall back
            n_loc = torch.tensor([vals.numel()], device=dev)
            ns = [torch.zeros_like(n_loc) for _ in range(c.world_size)]
            torch.distributed.all_gather(ns, n_loc)
            ns = [int(x) for x in ns]
            mx = max(ns + [1])
            pad_v = torch.full((mx,), -1, dtype=torch.int64, device=dev)
            pad_v[:vals.numel()] = vals
            pad_s = torch.zeros(mx, dtype=torch.int64, device=dev)
            pad_s[:vals.numel()] = seg
            gv = torch.zeros(c.world_size, mx, dtype=torch.int64, device=dev)
            gs = torch.zeros(c.world_size, mx, dtype=torch.int64, device=dev)
            c.all_gather(gv, pad_v)
            c.all_gather(gs, pad_s)
            gw = None
            if wts is not None:
                pad_w = torch.zeros(mx, device=dev)
                pad_w[:vals.numel()] = wts
                gw = torch.zeros(c.world_size, mx, device=dev)
                c.all_gather(gw, pad_w)
            owner, _ = var.locate(gv.clamp(min=0))
            mine = (owner == c.rank) & (gv >= 0)
            partial = torch.zeros(c.world_size, b_max(c, b), var.dim, device=dev)
            rows_all = torch.full_like(gv, -1)
            if bool(mine.any()):
                rws = var.local_rows(gv[mine])
                rows_all[mine] = rws
                src = torch.nonzero(mine)[:, 0]
                vec = var.weight[rws].float()
                if gw is not None:
                    vec = vec * gw[mine].unsqueeze(1)
                    ctx.w_own = gw[mine]
                partial.view(-1, var.dim).index_add_(0, src * partial.shape[1] + gs[mine], vec)
            recv = torch.zeros_like(partial)
            c.all_to_all(recv, partial)
            out = recv.sum(0)[:b]
            ctx.rows, ctx.seg_owner = rows_all, (gs, mine)
        if wts is None:
            cnt = lens.clamp(min=1).float().unsqueeze(1)
        else:      # weighted mean = weighted sum / sum of weights (tf.nn.embedding_lookup_sparse)
            cnt = torch.zeros(b, device=dev).index_add_(0, seg, wts).unsqueeze(1)
            cnt = torch.where(cnt == 0, torch.ones_like(cnt), cnt)
        if combiner == "mean":
            out = out / cnt
        ctx.var, ctx.seg, ctx.cnt, ctx.combiner, ctx.b = var, seg, cnt, combiner, b
        return out

    @staticmethod
    def backward(ctx, g):
        var, c = ctx.var, _c()
        g = g.float()
        if ctx.combiner == "mean":
            g = g / ctx.cnt
        if c.world_size == 1:
            rows, grads = ctx.rows, g[ctx.seg]
        else:
            gs, mine = ctx.seg_owner
            bm = b_max(c, ctx.b)
            gpad = torch.zeros(bm, var.dim, device=g.device)
            gpad[:ctx.b] = g
            gall = torch.zeros(c.world_size, bm, var.dim, device=g.device)
            c.all_gather(gall, gpad)
            src = torch.nonzero(mine)[:, 0]
            rows = ctx.rows[mine]
            grads = gall[src, gs[mine]]
        if ctx.w_own is not None and rows.numel():
            grads = grads * ctx.w_own.unsqueeze(1)
        if rows.numel():
            u, inv = torch.unique(rows, return_inverse=True)
            red = torch.zeros(u.numel(), var.dim, device=g.device).index_add_(0, inv, grads)
            if var.sparse_grad is None:
                var.sparse_grad = (u, red)
            else:
                var.sparse_grad = (torch.cat([var.sparse_grad[0], u]), torch.cat([var.sparse_grad[1], red]))
        return torch.zeros((), device=g.device), None, None, None, None, None


def b_max(c: Comm, b: int) -> int:
    t = torch.tensor([b], device=c.device)
    if c.world_size > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return int(t.item())


_anchor = None


def lookup_sparse(params, sp_ids, sp_weights=None, combiners: Union[str, Sequence[str], None] = None,
                  use_low_frequency_filter: bool = False):
    """sok.lookup_sparse(params, sp_ids, sp_weights=None, combiners=None) -> pooled embeddings
    [b, dim] per variable (reference lookup.py:543).  ``sp_weights``: per-id weights in the same layout
    as ``sp_ids`` (padded [b, H] or (values, row_lengths)); ``mean`` then divides by the sum of weights.
    A combiner (or list of combiners) passed in the third position is accepted as ``combiners``."""
    if isinstance(sp_weights, str) or (isinstance(sp_weights, (list, tuple)) and sp_weights
                                       and all(isinstance(x, str) for x in sp_weights)):
        sp_weights, combiners = None, sp_weights
    single = not isinstance(params, (list, tuple))
    ps = [params] if single else list(params)
    ids = [sp_ids] if single else list(sp_ids)
    ws = [None] * len(ps) if sp_weights is None or (isinstance(sp_weights, (list, tuple)) and not sp_weights) \
        else ([sp_weights] if single else list(sp_weights))
    combiners = combiners or "sum"
    cs = [combiners] * len(ps) if isinstance(combiners, str) else list(combiners)
    outs = []
    for p, i, w, cb in zip(ps, ids, ws, cs):
        vals, lens = _to_csr(i)
        wv = None
        if w is not None:
            if isinstance(w, (tuple, list)):
                wv = w[0].float().reshape(-1)
            else:
                wv = w.float()[i.to(torch.int64) >= 0] if w.dim() == 2 else w.float().reshape(-1)
            if wv.numel() != vals.numel():
                raise ValueError("sp_ids and sp_weights must have the same shape")
        anchor = torch.zeros((), device=p.device, requires_grad=True)
        outs.append(_Lookup.apply(anchor, p, vals, lens, cb, wv))
    return outs[0] if single else outs


def all2all_dense_embedding(param: Variable, indices: torch.Tensor) -> torch.Tensor:
    """dense (no pooling) lookup: [b, n] indices -> [b, n, dim]"""
    b, n = indices.shape
    out = lookup_sparse(param, indices.reshape(-1, 1), "sum")
    return out.view(b, n, param.dim)


# ----------------------------------------------------------------------------- optimizers
class OptimizerWrapper:
    """Applies sparse updates to SOK variables (sok.OptimizerWrapper over a tf optimizer): holds
    per-variable slot tensors, consumes ``var.sparse_grad``."""

    def __init__(self, optimizer: Optimizer_t = Optimizer_t.SGD, lr: float = 0.01, **hp):
        self.opt, self.lr, self.hp = optimizer, lr, hp
        self.step = 0

    def apply_gradients(self, variables: Sequence[Variable]):
        self.step += 1
        for v in variables:
            if v.sparse_grad is None or not v.trainable:
                v.sparse_grad = None
                continue
            rows, g = v.sparse_grad
            u, inv = torch.unique(rows, return_inverse=True)
            gg = torch.zeros(u.numel(), v.dim, device=g.device).index_add_(0, inv, g)
            for k in ("s0", "s1"):
                if k not in v.states:
                    v.states[k] = torch.full_like(v.weight, self.hp.get("initial_accu_value", 0.0)
                                                  if (k == "s0" and self.opt == Optimizer_t.AdaGrad) else 0.0,
                                                  dtype=torch.float32)
                elif v.states[k].shape[0] < v.weight.shape[0]:
                    ns = torch.zeros(v.weight.shape[0], v.dim, device=v.device)
                    ns[:v.states[k].shape[0]] = v.states[k]
                    v.states[k] = ns
            w = v.weight[u].float()
            a, b = v.states["s0"][u].clone(), v.states["s1"][u].clone()
            sparse_opt_reference(self.opt, w, a, b, gg, self.hp, self.lr, self.step)
            v.weight[u] = w.to(v.weight.dtype)
            v.states["s0"][u], v.states["s1"][u] = a, b
            v.last_touch[u.cpu()] = time.time()
            v.sparse_grad = None


def SGD(lr: float = 0.01):
    return OptimizerWrapper(Optimizer_t.SGD, lr)


def filter_variables(vars_):
    """split a variable list into (sok variables, others)"""
    sok_v = [v for v in vars_ if isinstance(v, Variable)]
    other = [v for v in vars_ if not isinstance(v, Variable)]
    return sok_v, other


# ----------------------------------------------------------------------------- dump / load
_MAGIC = b"SOKB200\0"


def _write(path, arr: np.ndarray, kind: int):
    with open(path, "wb") as f:
        f.write(_MAGIC + struct.pack("<iiqq", kind, arr.dtype.itemsize, arr.shape[0],
                                     arr.shape[1] if arr.ndim > 1 else 1))
        f.write(arr.tobytes())


def _read(path):
    raw = open(path, "rb").read()
    assert raw[:8] == _MAGIC, f"{path}: not a SOK dump"
    kind, isz, n, d = struct.unpack("<iiqq", raw[8:32])
    dt = {1: "<i8", 2: "<f4", 3: "<f4"}[kind]
    a = np.frombuffer(raw[32:], dtype=dt)
    return a.reshape(n, d) if kind != 1 else a


def dump(path: str, variables: Sequence[Variable], optimizer: Optional[OptimizerWrapper] = None):
    """<path>/<name>-key, <name>-weight (+ -slot0/-slot1), keys sorted, gathered on rank 0."""
    c = _c()
    os.makedirs(path, exist_ok=True)
    for v in variables:
        keys = v.global_keys()
        if isinstance(v, DynamicVariable):
            k, r = v.hash.dump()
            keys, w = k, v.weight[r.to(v.device)].float().cpu()
            sts = [v.states[s][r.to(v.device)].cpu() for s in ("s0", "s1") if s in v.states]
        else:
            w = v.weight.float().cpu()
            sts = [v.states[s].cpu() for s in ("s0", "s1") if s in v.states]
        parts = c.all_gather_object((keys, w, sts))
        if c.rank == 0:
            K = torch.cat([p[0] for p in parts])
            W = torch.cat([p[1] for p in parts])
            order = torch.argsort(K)
            _write(os.path.join(path, f"{v.name}-key"), K[order].numpy().astype("<i8"), 1)
            _write(os.path.join(path, f"{v.name}-weight"), W[order].numpy().astype("<f4"), 2)
            if optimizer is not None:
                for i in range(len(parts[0][2])):
                    S = torch.cat([p[2][i] for p in parts])
                    _write(os.path.join(path, f"{v.name}-slot{i}"), S[order].numpy().astype("<f4"), 3)
        c.barrier()


def load(path: str, variables: Sequence[Variable], optimizer: Optional[OptimizerWrapper] = None):
    for v in variables:
        K = torch.from_numpy(_read(os.path.join(path, f"{v.name}-key")).astype("int64"))
        W = torch.from_numpy(_read(os.path.join(path, f"{v.name}-weight")).copy())
        owner, _ = v.locate(K)
        m = owner == v.rank
        rows = v.local_rows(K[m])
        v.weight[rows] = W[m].to(v.device, v.weight.dtype)
        if optimizer is not None:
            for i, s in enumerate(("s0", "s1")):
                p = os.path.join(path, f"{v.name}-slot{i}")
                if os.path.exists(p):
                    S = torch.from_numpy(_read(p).copy())
                    if s not in v.states:
                        v.states[s] = torch.zeros_like(v.weight, dtype=torch.float32)
                    v.states[s][rows] = S[m].to(v.device)


def incremental_model_dump(variables: Sequence[DynamicVariable], time_threshold: float, path: str):
    """export only the rows touched after ``time_threshold`` (dump_load.py:1343-1500)."""
    os.makedirs(path, exist_ok=True)
    for v in variables:
        if isinstance(v, DynamicVariable):
            k, r = v.hash.dump()
        else:
            k, r = v.global_keys(), torch.arange(v.weight.shape[0])
        m = v.last_touch[r] > time_threshold
        _write(os.path.join(path, f"{v.name}-key"), k[m].numpy().astype("<i8"), 1)
        _write(os.path.join(path, f"{v.name}-weight"),
               v.weight[r[m].to(v.device)].float().cpu().numpy().astype("<f4"), 2)
