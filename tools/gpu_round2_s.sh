#!/bin/bash
# asinh / acosh region polynomials, atanh / pow table re-layout: tests, sweep, and dynamic pipe statistics (ncu) of
# every map kernel still below 80 % of the measured peak
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_transcendental_gpu.py tests/test_table_kernels_gpu.py -q -x > gpurun_out/gputest_trans_r02c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_trans_r02c.log; tail -3 gpurun_out/gputest_trans_r02c.log
python tools/sweep.py --only sasinh,sacosh,satanh,spow,ssinh,scosh,stan,dsin,dcos,dtan,dasin,dacos,datan,dsinh,dcosh,dtanh,dasinh,dacosh,datanh,datan2,dpow,dexp,dlog --reps 10 > gpurun_out/sweep_s.jsonl 2>&1
python - <<'P'
import json
for l in open('gpurun_out/sweep_s.jsonl'):
    if l.startswith('{'):
        r=json.loads(l); print("%-14s %.3f ms %.3f" % (r['symbol'], r['ms'], r['frac_of_measured_peak']))
P
SY=satanh,sasinh,sacosh,spow,ssinh,stan,dsin,dcos,dtan,dasin,dacos,datan,dsinh,dcosh,dtanh,dasinh,dacosh,datanh,datan2,dpow,dexp,dlog
python tools/prof_ops.py 24 $SY > gpurun_out/prof_s.log 2>&1 && \
timeout 900 ncu --section ComputeWorkloadAnalysis --section InstructionStats --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section SpeedOfLight \
  --clock-control none -k regex:k_map_flat -o gpurun_out/pipes_s -f python tools/prof_ops.py 24 $SY > gpurun_out/ncu_s.log 2>&1
tail -2 gpurun_out/ncu_s.log
