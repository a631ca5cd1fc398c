#!/usr/bin/env bash
# Build the host runtime (csrc/host/*.cpp) with AddressSanitizer+UBSan or ThreadSanitizer and run the
# native-code tests against it (HCTR_HOST_LIB overrides the library the package loads).
#   tools_dev/sanitize_host.sh asan|tsan
set -euo pipefail
cd "$(dirname "$0")/.."
MODE=${1:-asan}
OUT=/tmp/libhctr_host_${MODE}.so
if [ "$MODE" = asan ]; then
  FLAGS="-fsanitize=address,undefined -fno-omit-frame-pointer"; RT=$(gcc -print-file-name=libasan.so)
  export ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1
else
  FLAGS="-fsanitize=thread"; RT=$(gcc -print-file-name=libtsan.so)
  export TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0"
fi
g++ -O1 -g -std=c++17 -fPIC -pthread -fopenmp $FLAGS -shared -o "$OUT" hugectr_b200/csrc/host/*.cpp
LD_PRELOAD=$RT HCTR_HOST_LIB=$OUT python -m pytest -q -m "not gpu and not dist" \
  tests/test_norm_reader_cpu.py tests/test_datagen_cpu.py tests/test_criteo_preprocess_cpu.py \
  tests -k "norm or datagen or criteo or raw or hps or param_server or parquet or reader or csr" 2>&1 | tee /tmp/sanitize_${MODE}.log | tail -3
echo "sanitizer reports: $(grep -c 'ERROR: AddressSanitizer\|runtime error\|WARNING: ThreadSanitizer' /tmp/sanitize_${MODE}.log || true)"
