#!/bin/bash
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_mxfp8.py tests/test_readers_gpu.py tests/test_metrics_dist.py "tests/test_emu_ranks.py::test_emu_legacy_embeddings_fused_gpu" "tests/test_emu_ranks.py::test_emu_fused_benchmark_plan_gpu" "tests/test_layers_native_gpu.py::test_native_gru_matches_torch" -m gpu -q --timeout 40 > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
grep -v WARNING gpurun_out/h_pytest.log | tail -45 | cut -c1-230
