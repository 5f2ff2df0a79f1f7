#!/bin/bash
# tools/profile_round.sh -- everything profiles/ is refreshed from, run on the GPU box in one gpurun call:
# GPU tests, smoke, full sweep, bench (own arm + reference arm), then -- only after each command has exited 0
# without a profiler -- the ncu launch list of bench.py and --set full captures of the C2 and table kernels.
set -x
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/pytest_gpu.log 2>&1; echo pytest=$?
python __graft_entry__.py smoke > $o/smoke.log 2>&1; echo smoke=$?
python tools/sweep.py --out $o/sweep.jsonl > $o/sweep.log 2>&1; echo sweep=$?
python bench.py --impl reference > $o/bench_ref.json 2> $o/bench_ref.err; echo ref=$?
python bench.py > $o/bench.json 2> $o/bench.err; echo bench=$?
python bench.py --steps 2 --warmup 3 > $o/bench_short.json 2> $o/bench_short.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/launches_bench.csv \
      python bench.py --steps 2 --warmup 3 > $o/ncu_bench.log 2>&1; echo launches=$?
python tools/prof_c2.py 1 > $o/prof_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_map_flat -o $o/prof_c2_full -f \
      python tools/prof_c2.py 1 > $o/ncu_full.log 2>&1; echo c2full=$?
python tools/prof_ops.py 28 dexp,dlog,dtanh,dpow,spow,dsin,dcos,datan2 > $o/po_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_map_flat -o $o/prof_tab_full -f \
      python tools/prof_ops.py 28 dexp,dlog,dtanh,dpow,spow,dsin,dcos,datan2 > $o/ncu_tab.log 2>&1; echo tabfull=$?
