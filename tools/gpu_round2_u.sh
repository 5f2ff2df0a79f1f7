#!/bin/bash
# second pass over the double-float kernels: atan2 rewrite, 32-entry table for sinh / cosh, specialised roots; A/B of
# register caps / unroll, then dynamic pipe statistics of the new kernels
mkdir -p gpurun_out
V=numericals_b200/lib/variants
timeout 900 python -m pytest tests/test_transcendental_gpu.py tests/test_table_kernels_gpu.py -q -x > gpurun_out/gputest_trans_r02e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_trans_r02e.log; tail -3 gpurun_out/gputest_trans_r02e.log
T=dsin,dcos,dtan,dasin,dacos,datan,dsinh,dcosh,dtanh,dasinh,dacosh,datanh,datan2,dpow,dexp,dlog
{ echo "# default"; python tools/sweep.py --only $T --reps 10;
  for v in hm5 hm6 at4 atm0 sc4; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/sweep.py --only $T --reps 10; done; } > gpurun_out/ab_f64b_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_f64b_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:].split(' ')[0]
    elif line.startswith('{'):
        r=json.loads(line); tab.setdefault(r['symbol'],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items(): print("%-14s %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items())))
P
SY=dsin,dcos,dtan,dasin,dacos,datan,dsinh,dcosh,dtanh,dasinh,dacosh,datanh,datan2,dpow,satanh,spow
python tools/prof_ops.py 24 $SY > gpurun_out/prof_u.log 2>&1 && \
timeout 900 ncu --section ComputeWorkloadAnalysis --section InstructionStats --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section SpeedOfLight --section MemoryWorkloadAnalysis \
  --clock-control none -k regex:k_map_flat -o gpurun_out/pipes_u -f python tools/prof_ops.py 24 $SY > gpurun_out/ncu_u.log 2>&1
tail -2 gpurun_out/ncu_u.log
