#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
grep -v "WARNING" gpurun_out/c_pytest.log | tail -40
timeout 600 python bench.py --steps 20 --warmup 5 --no-standin --no-secondary --sustained-sec 0 --profile gpurun_out/c_prof_1gpu.txt > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
tail -c 700 gpurun_out/c_bench.json; tail -5 gpurun_out/c_bench.err
