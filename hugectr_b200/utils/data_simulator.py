"""Initializer simulators (C45): Uniform / Gaussian / VarianceScaling / Constant + sinusoidal
positional initialisation (HugeCTR/include/data_simulator.hpp, src/data_simulator.cu:32-172)."""
from __future__ import annotations

import math

import torch


class UniformDataSimulator:
    def __init__(self, lo: float, hi: float): self.lo, self.hi = lo, hi
    def fill(self, t: torch.Tensor, gen=None): return t.uniform_(self.lo, self.hi, generator=gen)


class GaussianDataSimulator:
    def __init__(self, mu: float, sigma: float, lo: float = -math.inf, hi: float = math.inf):
        self.mu, self.sigma, self.lo, self.hi = mu, sigma, lo, hi

    def fill(self, t, gen=None):
        t.normal_(self.mu, self.sigma, generator=gen)
        return t.clamp_(self.lo, self.hi) if math.isfinite(self.lo) or math.isfinite(self.hi) else t


class ConstantDataSimulator:
    def __init__(self, v: float): self.v = v
    def fill(self, t, gen=None): return t.fill_(self.v)


class VarianceScalingSimulator:
    """scale, mode in {fan_in, fan_out, fan_avg}, distribution in {uniform, norm}"""

    def __init__(self, scale: float, mode: str, distribution: str, fan_in: int, fan_out: int):
        n = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2.0}[mode]
        self.var = scale / max(n, 1.0)
        self.distribution = distribution

    def fill(self, t, gen=None):
        if self.distribution == "uniform":
            lim = math.sqrt(3.0 * self.var)
            return t.uniform_(-lim, lim, generator=gen)
        return t.normal_(0.0, math.sqrt(self.var), generator=gen)


def sinusoidal_init(max_sequence_len: int, ev_size: int) -> torch.Tensor:
    """positional table used by InitParams(Sinusoidal) (data_simulator.cu:39)"""
    pos = torch.arange(max_sequence_len, dtype=torch.float32).unsqueeze(1)
    i = torch.arange(ev_size, dtype=torch.float32).unsqueeze(0)
    angle = pos / torch.pow(10000.0, (2 * torch.div(i, 2, rounding_mode="floor")) / ev_size)
    out = torch.zeros(max_sequence_len, ev_size)
    out[:, 0::2] = torch.sin(angle[:, 0::2])
    out[:, 1::2] = torch.cos(angle[:, 1::2])
    return out
