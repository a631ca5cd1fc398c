"""Evaluation metrics: AUC (exact, distributed), AverageLoss, HitRate, NDCG, SMAPE.

Reference: HugeCTR/src/metrics.cu:34-2155, HugeCTR/include/metrics.hpp:44-554.  Interface kept:
``local_reduce(raw)`` per eval batch, ``global_reduce()``, ``finalize_metric()``.

AUC: exact trapezoid AUC with tie handling (device sort + cumulative TP/FP + trapz).  Multi-GPU:
every rank's (pred, label) pairs of the evaluation round are all-gathered (a few MB per round) and the
exact AUC is computed redundantly on each rank -- the reference's histogram -> pivots -> all-to-all ->
local sort -> halo pipeline (metrics.cu:1017-1240) exists to bound memory on 16 GB V100s and is not
needed at B200 capacities.  Multi-class AUC = unweighted macro average of per-class AUCs.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .enums import MetricsRawType, MetricsType


def auc_exact(pred: torch.Tensor, label: torch.Tensor) -> float:
    """Trapezoid AUC with ties, any device."""
    p = pred.reshape(-1).float()
    y = (label.reshape(-1).float() > 0.5).float()
    n = p.numel()
    if n == 0:
        return 0.0
    order = torch.argsort(p, descending=True)
    p, y = p[order], y[order]
    # last index of each run of equal preds
    is_last = torch.ones(n, dtype=torch.bool, device=p.device)
    is_last[:-1] = p[1:] != p[:-1]
    tps = torch.cumsum(y, 0)[is_last].double()
    fps = torch.cumsum(1 - y, 0)[is_last].double()
    P, N = tps[-1], fps[-1]
    if P == 0 or N == 0:
        return 0.0  # undefined; reference returns 0 with a warning
    z = torch.zeros(1, dtype=torch.double, device=p.device)
    tps = torch.cat([z, tps])
    fps = torch.cat([z, fps])
    area = torch.trapz(tps, fps)
    return float(area / (P * N))


class Metric:
    name = "metric"

    def __init__(self, comm, num_classes: int = 1):
        self.comm = comm
        self.num_classes = num_classes
        self.reset()

    def reset(self):
        pass

    def set_current_batch_size(self, n: int):
        self.current_batch = n

    def local_reduce(self, raw: Dict[MetricsRawType, torch.Tensor]):
        raise NotImplementedError

    def global_reduce(self):
        pass

    def finalize_metric(self) -> float:
        raise NotImplementedError


class AUC(Metric):
    name = "AUC"

    def reset(self):
        self.preds: List[torch.Tensor] = []
        self.labels: List[torch.Tensor] = []
        self.per_class: List[float] = []

    def local_reduce(self, raw):
        self.preds.append(raw[MetricsRawType.Pred].detach().float().reshape(-1, self.num_classes).clone())
        self.labels.append(raw[MetricsRawType.Label].detach().float().reshape(-1, self.num_classes).clone())

    def _gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.comm is None or self.comm.world_size == 1:
            return t
        # equal-sized shards per rank (eval batches are split evenly)
        out = torch.empty(self.comm.world_size * t.numel(), dtype=t.dtype, device=t.device)
        self.comm.all_gather(out, t.contiguous())
        return out.view(self.comm.world_size * t.shape[0], t.shape[-1])

    def finalize_metric(self) -> float:
        if not self.preds:
            return 0.0
        p = self._gather(torch.cat(self.preds))
        y = self._gather(torch.cat(self.labels))
        self.per_class = [auc_exact(p[:, c], y[:, c]) for c in range(self.num_classes)]
        self.preds, self.labels = [], []
        return sum(self.per_class) / len(self.per_class)


class AverageLoss(Metric):
    name = "AverageLoss"

    def reset(self):
        self.total = 0.0
        self.batches = 0

    def local_reduce(self, raw):
        v = raw[MetricsRawType.Loss].detach().float().sum()
        if self.comm is not None and self.comm.world_size > 1:
            v = v.clone()
            self.comm.all_reduce(v)
            v = v / self.comm.world_size
        self.total += float(v)
        self.batches += 1

    def finalize_metric(self) -> float:
        r = self.total / max(1, self.batches)
        self.reset()
        return r


class HitRate(Metric):
    """hits / checked where checked = pred > 0.8 (metrics.cu:1748-1758)."""
    name = "HitRate"

    def reset(self):
        self.checked = 0.0
        self.hits = 0.0

    def local_reduce(self, raw):
        p = raw[MetricsRawType.Pred].detach().float().reshape(-1)
        y = raw[MetricsRawType.Label].detach().float().reshape(-1)
        m = p > 0.8
        c = torch.stack([m.sum().float(), (m & (y == 1.0)).sum().float()])
        if self.comm is not None and self.comm.world_size > 1:
            self.comm.all_reduce(c)
        self.checked += float(c[0])
        self.hits += float(c[1])

    def finalize_metric(self) -> float:
        r = self.hits / self.checked if self.checked > 0 else 0.0
        self.reset()
        return r


class SMAPE(Metric):
    """mean(|p-y| / ((p+y)/2)) (metrics.cu:1882-1889)."""
    name = "SMAPE"

    def reset(self):
        self.err = 0.0
        self.n = 0.0

    def local_reduce(self, raw):
        p = raw[MetricsRawType.Pred].detach().float().reshape(-1)
        y = raw[MetricsRawType.Label].detach().float().reshape(-1)
        c = torch.stack([((p - y).abs() / ((p + y) / 2)).sum(),
                         torch.tensor(float(p.numel()), device=p.device)])
        if self.comm is not None and self.comm.world_size > 1:
            self.comm.all_reduce(c)
        self.err += float(c[0])
        self.n += float(c[1])

    def finalize_metric(self) -> float:
        r = self.err / self.n if self.n > 0 else 0.0
        self.reset()
        return r


class NDCG(AUC):
    """DCG of labels sorted by pred / ideal DCG (metrics.cu:1656-1706)."""
    name = "NDCG"

    def finalize_metric(self) -> float:
        if not self.preds:
            return 0.0
        p = self._gather(torch.cat(self.preds)).reshape(-1)
        y = self._gather(torch.cat(self.labels)).reshape(-1)
        self.preds, self.labels = [], []
        n = p.numel()
        disc = 1.0 / torch.log2(torch.arange(n, device=p.device, dtype=torch.double) + 2.0)
        dcg = (y[torch.argsort(p, descending=True)].double() * disc).sum()
        idcg = (torch.sort(y, descending=True).values.double() * disc).sum()
        return float(dcg / idcg) if idcg > 0 else 0.0


def create_metric(kind: MetricsType, comm, num_classes: int = 1) -> Metric:
    return {MetricsType.AUC: AUC, MetricsType.AverageLoss: AverageLoss,
            MetricsType.HitRate: HitRate, MetricsType.NDCG: NDCG,
            MetricsType.SMAPE: SMAPE}[kind](comm, num_classes)
