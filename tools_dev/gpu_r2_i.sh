#!/bin/bash
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_mxfp8.py -m gpu -q --timeout 30 > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
grep -v WARNING gpurun_out/i_pytest.log | tail -8 | cut -c1-200
timeout 40 python tools_dev/bench_mxfp8.py > gpurun_out/i_mxfp8.jsonl 2> gpurun_out/i_mxfp8.err; echo "mx rc=$?"
cat gpurun_out/i_mxfp8.jsonl | cut -c1-420; tail -3 gpurun_out/i_mxfp8.err
