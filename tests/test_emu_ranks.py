"""Multi-rank paths on ONE device through the in-process fabric (hugectr_b200/parallel/emu.py): the same
entry points as tests/test_dist.py (torchrun, real GPUs), run as N threads.  On a GPU the collections are in
FUSED mode: owner-side kernels read the other ranks' key slabs, write their output slabs and read their
gradient slabs through "peer" pointers (ordinary device pointers here), the device barrier and the two-shot
all-reduce kernels rendezvous across the ranks' streams -- this is the driver-visible proof for the peer
kernels on a one-GPU box.  On CPU the collective path is exercised without process spawns."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import dist_worker as W  # noqa: E402
from hugectr_b200.parallel.emu import run_ranks  # noqa: E402

CPU = torch.device("cpu")


@pytest.mark.parametrize("plan", ["mixed", "column"])
def test_emu_ebc_collective_cpu(plan):
    run_ranks(2 if plan == "column" else 3, lambda c: W.run_ebc(plan, False, comm=c), device=CPU)


def test_emu_equiv_cpu():
    run_ranks(2, lambda c: W.run_equiv("adagrad", comm=c), device=CPU)


def test_emu_fuzz_cpu():
    for seed in (11, 12):
        run_ranks(2, lambda c: W.run_fuzz(seed, comm=c), device=CPU)


@pytest.fixture
def shard_split_env():
    old = os.environ.get("HCTR_SHARD_SPLIT")
    os.environ["HCTR_SHARD_SPLIT"] = "1"
    yield
    if old is None:
        os.environ.pop("HCTR_SHARD_SPLIT", None)
    else:
        os.environ["HCTR_SHARD_SPLIT"] = old


def test_emu_shard_split_cpu(shard_split_env):
    run_ranks(2, lambda c: W.run_ebc("mixed", False, comm=c), device=CPU)
    run_ranks(2, lambda c: W.run_equiv("sgd", comm=c), device=CPU)


# ----------------------------------------------------------------------------- GPU: fused peer kernels
@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("plan", ["mixed", "column"])
def test_emu_ebc_fused_gpu(plan, world):
    run_ranks(world, lambda c: W.run_ebc(plan, True, comm=c))


@pytest.mark.gpu
def test_emu_ebc_fused_shard_split_gpu(shard_split_env):
    run_ranks(4, lambda c: W.run_ebc("mixed", True, comm=c))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_emu_fuzz_fused_gpu(seed):
    run_ranks(4, lambda c: W.run_fuzz(seed, comm=c))


@pytest.mark.gpu
def test_emu_p2p_allreduce_gpu():
    run_ranks(4, lambda c: W.run_allreduce(comm=c))


@pytest.mark.gpu
@pytest.mark.parametrize("opt", ["sgd", "adagrad"])
def test_emu_model_equals_single_gpu(opt):
    run_ranks(4, lambda c: W.run_equiv(opt, comm=c))


@pytest.mark.gpu
def test_emu_model_equals_single_shard_split_gpu(shard_split_env):
    run_ranks(4, lambda c: W.run_equiv("adagrad", comm=c))


@pytest.mark.gpu
def test_emu_model_overlap_paths_gpu():
    run_ranks(2, lambda c: W.run_model(comm=c))


@pytest.mark.gpu
def test_emu_dynamic_tables_fused_gpu():
    """dynamic (hashed) tables on the fused path: owners translate the keys that landed in their inbox"""
    run_ranks(4, lambda c: W.run_dynamic(comm=c))


def test_emu_dynamic_tables_cpu():
    run_ranks(2, lambda c: W.run_dynamic(comm=c), device=CPU)


def test_emu_legacy_embeddings_cpu():
    run_ranks(2, lambda c: W.run_legacy(comm=c), device=CPU)


@pytest.mark.gpu
def test_emu_legacy_embeddings_fused_gpu():
    """Distributed / Localized hash embeddings on the fused path (peer-store key / gradient exchange,
    one-kernel ownership filter + hash translation): replicas stay identical, loss finite"""
    run_ranks(4, lambda c: W.run_legacy(comm=c))
