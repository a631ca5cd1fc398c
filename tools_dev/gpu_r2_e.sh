#!/bin/bash
# N-GPU bench: our arms (+profile) and the reference arm
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --profile gpurun_out/e_prof_${N}gpu.txt > gpurun_out/e_bench_${N}gpu.json 2> gpurun_out/e_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 2600 gpurun_out/e_bench_${N}gpu.json; grep -v "WARNING\|^$\|OMP_NUM\|\*\*\*\*\|profiler\|_warn_once" gpurun_out/e_bench_${N}gpu.err | tail -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/e_ref_${N}gpu.json 2> gpurun_out/e_ref_${N}gpu.err; echo "ref rc=$?"
tail -c 1800 gpurun_out/e_ref_${N}gpu.json; grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/e_ref_${N}gpu.err | tail -c 600
