#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reduce_gpu.py tests/test_bmas_gpu.py -q -x > gpurun_out/gputest_reduce_r02.log 2>&1; tail -1 gpurun_out/gputest_reduce_r02.log
python tools/sweep.py --only i8sum,i8hmax,u8hmax,i16hmax,u16hmax,ssum,i8dot --reps 10 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print(r['symbol'], r['ms'], r['frac_of_measured_peak'])"
