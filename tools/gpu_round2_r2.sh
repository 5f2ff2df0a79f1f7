#!/bin/bash
# software-pipelined k_reduce_flat (next tile's loads before the current tile's fold), same box A/B over every full reduction
mkdir -p gpurun_out
V=numericals_b200/lib/variants
timeout 600 python -m pytest tests/test_reduce_gpu.py -q -x > gpurun_out/gputest_reduce_default.log 2>&1; tail -1 gpurun_out/gputest_reduce_default.log
NUK_LIB=$V/libnuk_rpf.so timeout 600 python -m pytest tests/test_reduce_gpu.py tests/test_configs_gpu.py -q -x > gpurun_out/gputest_reduce_rpf.log 2>&1; tail -1 gpurun_out/gputest_reduce_rpf.log
{ for rep in 1 2; do echo "# default$rep"; python tools/sweep.py --only sum,hmax --reps 10; echo "# rpf$rep"; NUK_LIB=$V/libnuk_rpf.so python tools/sweep.py --only sum,hmax --reps 10; done; } > gpurun_out/ab_reduce_prefetch_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_reduce_prefetch_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:].split(' ')[0]
    elif line.startswith('{'):
        r=json.loads(line); tab.setdefault(r['symbol'],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items(): print("%-14s %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items())))
P
