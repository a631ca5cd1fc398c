#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -30 gpurun_out/b_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-standin --no-secondary --sustained-sec 0 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/b_bench.json; tail -5 gpurun_out/b_bench.err
HCTR_CONCAT_ALIAS=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-standin --no-secondary --sustained-sec 0 > gpurun_out/b_bench_alias.json 2> gpurun_out/b_bench_alias.err; echo "alias rc=$?"
tail -c 600 gpurun_out/b_bench_alias.json; tail -5 gpurun_out/b_bench_alias.err
