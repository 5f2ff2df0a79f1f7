#!/bin/bash
# pass-1 grid size of the full reductions for the short (8-bit, 268 MB) problems
mkdir -p gpurun_out
{ for m in 16 4 8 12 24 32; do echo "# mult$m"; NUK_REDUCE_GRID_MULT=$m python tools/sweep.py --only i8sum,i8hmax,u8hmax,i16hmax,ssum --reps 10; done; } > gpurun_out/ab_reduce_grid_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_reduce_grid_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:].split(' ')[0]
    elif line.startswith('{'):
        r=json.loads(line); tab.setdefault(r['symbol'],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items(): print("%-14s %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items())))
P
