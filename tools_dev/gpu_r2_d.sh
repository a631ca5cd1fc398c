#!/bin/bash
# 2-GPU check: real multi-GPU tests (NCCL + IPC heap + fused dispatch/push kernels), bench at N=2 (all arms), reference N=2
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_dist.py -m gpu -q --timeout 400 > gpurun_out/d_pytest_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest_${N}gpu.log
grep -v "WARNING" gpurun_out/d_pytest_${N}gpu.log | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --profile gpurun_out/d_prof_${N}gpu.txt > gpurun_out/d_bench_${N}gpu.json 2> gpurun_out/d_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/d_bench_${N}gpu.json; grep -v "WARNING\|^$" gpurun_out/d_bench_${N}gpu.err | tail -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/d_ref_${N}gpu.json 2> gpurun_out/d_ref_${N}gpu.err; echo "ref rc=$?"
tail -c 1800 gpurun_out/d_ref_${N}gpu.json; tail -c 800 gpurun_out/d_ref_${N}gpu.err
