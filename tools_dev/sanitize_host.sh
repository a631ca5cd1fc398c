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
  OMP=-fopenmp
else
  # ThreadSanitizer does not model libgomp's fork / join barriers (every `omp parallel` region reads as a race between
  # the master's set-up and the workers): the TSan build compiles the pragmas away, so what is checked is the
  # runtime's OWN threading -- reader worker threads, ring hand-over, parameter-server shards, the exec watchdog.
  # Reports go to files (pytest captures stderr of passing tests).
  FLAGS="-fsanitize=thread -Wno-unknown-pragmas"; RT=$(gcc -print-file-name=libtsan.so)
  rm -f /tmp/tsan_report.*
  export TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0 log_path=/tmp/tsan_report"
  OMP=
fi
g++ -O1 -g -std=c++17 -fPIC -pthread $OMP $FLAGS -shared -o "$OUT" hugectr_b200/csrc/host/*.cpp
LD_PRELOAD=$RT HCTR_HOST_LIB=$OUT python -m pytest -q -m "not gpu and not dist" \
  tests/test_norm_reader_cpu.py tests/test_datagen_cpu.py tests/test_criteo_preprocess_cpu.py \
  tests -k "norm or datagen or criteo or raw or hps or param_server or parquet or reader or csr or watchdog" 2>&1 | tee /tmp/sanitize_${MODE}.log | tail -3
echo "sanitizer reports: $(cat /tmp/sanitize_${MODE}.log /tmp/tsan_report.* 2>/dev/null | grep -c 'ERROR: AddressSanitizer\|runtime error\|WARNING: ThreadSanitizer' || true)"
if [ "$MODE" = tsan ]; then
  echo "in libhctr_host: $(cat /tmp/tsan_report.* 2>/dev/null | grep -c 'libhctr_host_tsan.so' || true) frames"
fi
