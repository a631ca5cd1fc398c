"""Distributed exact AUC (histogram -> ranges -> all-to-all -> local sort) against the single-process
trapezoid AUC, with ties, unequal per-rank counts and an empty rank; CPU (emulated ranks / torch ops) and
GPU (csrc/metrics.cu kernels)."""
import pytest
import torch

from hugectr_b200.metrics import AUC, auc_distributed, auc_exact
from hugectr_b200.enums import MetricsRawType
from hugectr_b200.parallel.emu import run_ranks


def _data(dev):
    g = torch.Generator().manual_seed(3)
    N = [5000, 3000, 0, 7001]
    preds = [(torch.sigmoid(torch.randn(n, generator=g) * 2) * 50).round() / 50 if i % 2 == 0
             else torch.sigmoid(torch.randn(n, generator=g)) for i, n in enumerate(N)]
    preds[1][:100] = -1.5              # values outside [0, 1] and negative keys
    labels = [(torch.rand(n, generator=g) < 0.3).float() for n in N]
    return [p.to(dev) for p in preds], [y.to(dev) for y in labels]


def _check(dev):
    preds, labels = _data(dev)
    exp = auc_exact(torch.cat(preds).cpu(), torch.cat(labels).cpu())
    res = run_ranks(4, lambda c: auc_distributed(preds[c.rank], labels[c.rank], c), device=dev)
    assert all(abs(r - exp) < 1e-9 for r in res), (exp, res)

    def metric(c):
        m = AUC(c, 1)
        for lo in range(0, preds[c.rank].numel(), 1024):        # several eval batches per rank
            m.local_reduce({MetricsRawType.Pred: preds[c.rank][lo:lo + 1024],
                            MetricsRawType.Label: labels[c.rank][lo:lo + 1024]})
        return m.finalize_metric()
    res = run_ranks(4, metric, device=dev)
    assert all(abs(r - exp) < 1e-9 for r in res), (exp, res)


def test_auc_distributed_cpu():
    _check(torch.device("cpu"))


@pytest.mark.gpu
def test_auc_distributed_gpu():
    _check(torch.device("cuda"))


@pytest.mark.gpu
def test_auc_distributed_large_gpu():
    """1 M pairs per rank on 4 emulated ranks; degenerate distributions (all equal) stay exact"""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    preds = [torch.sigmoid(torch.randn(1 << 20, generator=g, device=dev) - 2) for _ in range(4)]
    labels = [(torch.rand(1 << 20, generator=g, device=dev) < preds[i]).float() for i in range(4)]
    exp = auc_exact(torch.cat(preds), torch.cat(labels))
    res = run_ranks(4, lambda c: auc_distributed(preds[c.rank], labels[c.rank], c), device=dev)
    assert all(abs(r - exp) < 1e-7 for r in res), (exp, res)
    same = [torch.full((1000,), 0.25, device=dev) for _ in range(4)]
    lab = [(torch.arange(1000, device=dev) % 3 == 0).float() for _ in range(4)]
    res = run_ranks(4, lambda c: auc_distributed(same[c.rank], lab[c.rank], c), device=dev)
    assert all(abs(r - 0.5) < 1e-12 for r in res), res
