#!/bin/bash
# double-float asinh / acosh region polynomials, select-free atan, specialised root in asin / acos, atanh quotient,
# whole-tile batch for sinh / cosh: tests + A/B sweeps inside one call
mkdir -p gpurun_out
V=numericals_b200/lib/variants
timeout 900 python -m pytest tests/test_transcendental_gpu.py tests/test_table_kernels_gpu.py -q -x > gpurun_out/gputest_trans_r02d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_trans_r02d.log; tail -3 gpurun_out/gputest_trans_r02d.log
T=dsin,dcos,dtan,dasin,dacos,datan,dsinh,dcosh,dtanh,dasinh,dacosh,datanh,datan2,dpow,dexp,dlog
{ echo "# default"; python tools/sweep.py --only $T --reps 10;
  for v in hypnb ihyp2 invb invb4; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/sweep.py --only $T --reps 10; done; } > gpurun_out/ab_f64_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_f64_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:].split(' ')[0]
    elif line.startswith('{'):
        r=json.loads(line); tab.setdefault(r['symbol'],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items(): print("%-14s %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items())))
P
