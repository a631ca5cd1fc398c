"""CSR sparse tensor (value / row offset / nnz) of the legacy data path (C44: SparseTensor<T>,
SparseTensor23) and conversions to / from the padded feature-major layout used by the kernels."""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class SparseTensor:
    values: torch.Tensor        # [nnz]
    row_offsets: torch.Tensor   # [rows + 1]

    @property
    def nnz(self) -> int:
        return int(self.row_offsets[-1])

    @property
    def rows(self) -> int:
        return self.row_offsets.numel() - 1

    @staticmethod
    def from_padded(keys: torch.Tensor) -> "SparseTensor":
        """keys [rows, H] padded with -1"""
        m = keys >= 0
        lens = m.sum(1)
        off = torch.zeros(keys.shape[0] + 1, dtype=torch.int64, device=keys.device)
        off[1:] = torch.cumsum(lens, 0)
        return SparseTensor(keys[m], off)

    def to_padded(self, max_nnz: int) -> torch.Tensor:
        rows = self.rows
        out = torch.full((rows, max_nnz), -1, dtype=self.values.dtype, device=self.values.device)
        lens = (self.row_offsets[1:] - self.row_offsets[:-1]).clamp(max=max_nnz)
        pos = torch.arange(max_nnz, device=out.device).unsqueeze(0)
        m = pos < lens.unsqueeze(1)
        src = (self.row_offsets[:-1].unsqueeze(1) + pos)[m]
        out[m] = self.values[src]
        return out
