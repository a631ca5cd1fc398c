"""embedding_workspace_calculator: workspace_size_per_gpu_in_mb for legacy SparseEmbedding.

max_vocabulary_size_per_gpu = workspace_MB * 2^20 / ((1 + n_opt_states) * 4 * vec)  (model.cpp:186-196)
-> the inverse, with the reference tool's safety factor (tools/embedding_workspace_calculator).
"""
from __future__ import annotations

import math

from ..enums import OPT_STATES_PER_WEIGHT, Embedding_t, Optimizer_t, Update_t


def calculate(slot_size_array, vec_size: int, optimizer: Optimizer_t = Optimizer_t.Adam,
              update_type: Update_t = Update_t.Global, num_gpus: int = 1,
              embedding_type: Embedding_t = Embedding_t.DistributedSlotSparseEmbeddingHash,
              load_factor: float = 0.75) -> int:
    ns = OPT_STATES_PER_WEIGHT[optimizer]
    if optimizer == Optimizer_t.Adam and update_type == Update_t.LazyGlobal:
        ns += 1
    if embedding_type == Embedding_t.DistributedSlotSparseEmbeddingHash:
        rows = math.ceil(sum(slot_size_array) / num_gpus)
    else:
        per = [0] * num_gpus
        for i, s in enumerate(slot_size_array):
            per[i % num_gpus] += s
        rows = max(per)
    rows = math.ceil(rows / load_factor)
    return math.ceil(rows * (1 + ns) * 4 * vec_size / (1 << 20))
