#!/bin/bash
# tools/sanitize.sh -- compute-sanitizer passes over the GPU suite (the reference has no race
# detector; this is the replacement's): memcheck over the parity tests of every kernel family, racecheck over the
# kernels that use shared memory (reductions, dot / arg-reductions, table-staging maps, TMA column reduction).
# Writes gpurun_out/sanitizer_r02.txt; the large-array tests (configs, multi-GB sweeps) are left out: the tools
# slow every kernel 10-100x.
mkdir -p gpurun_out
out=gpurun_out/sanitizer_r02.txt
: > $out
run() {  # <tool> <timeout> <pytest args...>
  tool=$1; to=$2; shift 2
  echo "=== compute-sanitizer --tool $tool :: pytest $*" >> $out
  timeout $to compute-sanitizer --tool $tool --error-exitcode 86 --launch-timeout 0 python -m pytest "$@" -q -x -p no:cacheprovider >> $out 2>&1
  echo "=== exit code $? (86 = the tool reported an error)" >> $out
}
run memcheck 1500 tests/test_bmas_gpu.py tests/test_nd_gpu.py tests/test_reduce_gpu.py tests/test_next_rows_gpu.py tests/test_runtime_gpu.py
run racecheck 900 tests/test_reduce_gpu.py tests/test_next_rows_gpu.py tests/test_table_kernels_gpu.py
grep -E "^===|passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $out
