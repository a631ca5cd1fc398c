"""Host thread pool + per-replica CPU resources (C47: src/thread_pool.cpp, src/cpu_resource.cpp).
``HCTR_DEFAULT_CONCURRENCY`` sizes the default pool."""
from __future__ import annotations

import os
from concurrent.futures import Future, ThreadPoolExecutor

import torch


class ThreadPool:
    _inst = None

    def __init__(self, n: int = 0):
        n = n or int(os.environ.get("HCTR_DEFAULT_CONCURRENCY", "0")) or min(32, os.cpu_count() or 4)
        self.n = n
        self.ex = ThreadPoolExecutor(max_workers=n, thread_name_prefix="hctr")

    @classmethod
    def get(cls) -> "ThreadPool":
        if cls._inst is None:
            cls._inst = ThreadPool()
        return cls._inst

    def submit(self, fn, *a, **kw) -> Future:
        return self.ex.submit(fn, *a, **kw)

    def await_idle(self, futures):
        return [f.result() for f in futures]


class CPUResource:
    """replica-uniform and replica-variant host generators (same seed on every replica for weights)"""

    def __init__(self, replica_uniform_seed: int, replica_variant_seeds):
        self.uniform = torch.Generator().manual_seed(int(replica_uniform_seed))
        self.variant = [torch.Generator().manual_seed(int(s)) for s in replica_variant_seeds]

    def get_replica_uniform_generator(self): return self.uniform
    def get_replica_variant_generator(self, i: int): return self.variant[i]
