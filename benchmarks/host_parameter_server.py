#!/usr/bin/env python
"""Host parameter server throughput (the host tier of the embedding training cache): key -> row resolution with row
creation, row gather (pull) and write-through (push) per step-sized key batch.  No GPU involved.

  python benchmarks/host_parameter_server.py [--rows 4000000] [--ev 128] [--batch 200000]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from hugectr_b200.cache.hps import HostParameterServer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_000_000)
    ap.add_argument("--ev", type=int, default=128)
    ap.add_argument("--batch", type=int, default=200_000, help="distinct keys per step")
    ap.add_argument("--states", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    ps = HostParameterServer(a.ev, a.states, capacity_rows=a.rows + 1024, seed=1)
    g = torch.Generator().manual_seed(0)
    t0 = time.perf_counter()
    for lo in range(0, a.rows, 1 << 20):                      # populate (row creation + init)
        k = torch.arange(lo, min(a.rows, lo + (1 << 20)), dtype=torch.int64) * 2654435761 % (1 << 40)
        ps._rows(k, create=True)
    fill_s = time.perf_counter() - t0
    pull = push = 0.0
    for _ in range(a.steps):
        idx = torch.randint(0, a.rows, (a.batch,), generator=g)
        k = torch.unique(idx.to(torch.int64) * 2654435761 % (1 << 40))
        t0 = time.perf_counter()
        w, sts = ps.pull(k)
        pull += time.perf_counter() - t0
        t0 = time.perf_counter()
        ps.push(k, w, sts)
        push += time.perf_counter() - t0
        n = k.numel()
    row_bytes = a.ev * 4 * (1 + a.states)
    print(json.dumps({"rows": a.rows, "ev": a.ev, "states": a.states, "keys_per_step": int(n),
                      "populate_Mrows_per_s": round(a.rows / fill_s / 1e6, 2),
                      "pull_Mkeys_per_s": round(n * a.steps / pull / 1e6, 2),
                      "pull_GB_per_s": round(n * a.steps * row_bytes / pull / 1e9, 2),
                      "push_Mkeys_per_s": round(n * a.steps / push / 1e6, 2),
                      "push_GB_per_s": round(n * a.steps * row_bytes / push / 1e9, 2),
                      "host_cores": os.cpu_count(), "size": ps.size()}))


if __name__ == "__main__":
    main()
