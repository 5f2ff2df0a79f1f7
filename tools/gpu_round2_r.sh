#!/bin/bash
# A/B of the single-float asin / acos / atan / atan2 / sinh / cosh rewrites (select-free regions, pure single float)
mkdir -p gpurun_out
V=numericals_b200/lib/variants
T=ssinh,scosh,satan,sasin,sacos,satan2,stan,stanh
timeout 900 python -m pytest tests/test_transcendental_gpu.py tests/test_table_kernels_gpu.py -q -x > gpurun_out/gputest_trans_r02b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_trans_r02b.log; tail -3 gpurun_out/gputest_trans_r02b.log
{ echo "# default (warp split for asin acos atan atan2, pack split for sinh cosh)"; python tools/sweep.py --only $T --reps 10;
  for v in hyp2 at1 at0 u8; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/sweep.py --only $T --reps 10; done; } > gpurun_out/ab_invtrig_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_invtrig_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:].split(' ')[0]
    elif line.startswith('{'):
        r=json.loads(line); tab.setdefault(r['symbol'],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items(): print("%-14s %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items())))
P
