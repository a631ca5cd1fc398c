#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_readers_gpu.py tests/test_mxfp8.py -q --timeout 50 -x > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
grep -v WARNING gpurun_out/g_pytest.log | tail -25
HCTR_SYNTH_POOL=8 timeout 60 python bench.py --steps 10 --warmup 3 --no-standin --no-secondary --sustained-sec 0.3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
tail -c 900 gpurun_out/g_bench.json; grep "bench \|Error\|error" gpurun_out/g_bench.err | tail
