"""torchrun worker of tests/test_bench_contract_cpu.py: attempt 0 wedges (rank 1 never enters the collective
and holds the GIL, rank 0 blocks inside it); the bench's headline watchdog re-executes every rank with the next
schedule, the new images meet on a fresh port and finish."""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    attempt = int(os.environ.get("HCTR_BENCH_ATTEMPT", "0"))
    rank = int(os.environ["RANK"])
    disarm = bench.arm_headline_watchdog(float(os.environ.get("WD_SECONDS", "3")), attempt, rank,
                                         argv=[sys.executable, os.path.abspath(__file__)])
    dist.init_process_group("gloo")
    if attempt == 0:
        if rank == 1:
            ctypes.PyDLL(None).sleep(600)          # never arrives; keeps the GIL
        t = torch.ones(1)
        dist.all_reduce(t)                          # rank 0 waits for a peer that never comes
        print("NOT REACHED", flush=True)
        return 1
    t = torch.full((1,), float(rank + 1))
    dist.all_reduce(t)
    disarm()
    if rank == 0:
        print(f"OK attempt={attempt} sum={t.item()} overlap_off={os.environ.get('HCTR_DISABLE_OVERLAP')} "
              f"agent_store={os.environ.get('TORCHELASTIC_USE_AGENT_STORE')}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
