#!/bin/bash
# round-2 GPU check A: tests, bench (all arms), reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -3 gpurun_out/a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile gpurun_out/a_prof_1gpu.txt > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/a_ref.json 2> gpurun_out/a_ref.err; echo "ref rc=$?"
tail -c 2500 gpurun_out/a_ref.json; tail -c 1500 gpurun_out/a_ref.err
