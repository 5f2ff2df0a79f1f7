#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_transcendental_gpu.py tests/test_table_kernels_gpu.py -q -x > gpurun_out/gputest_trans_r02j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_trans_r02j.log; tail -2 gpurun_out/gputest_trans_r02j.log
for c in 0 1; do python tools/sweep.py --only dtanh,dsinh,dcosh,dexp --reps 10 --cooldown $c 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print(r['symbol'], r['ms'], r['frac_of_measured_peak'], r.get('cooldown_s',0))"; done
