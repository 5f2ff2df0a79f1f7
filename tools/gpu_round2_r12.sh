#!/bin/bash
# double asin / acos: register cap at 3 packs per thread, same box
mkdir -p gpurun_out
V=numericals_b200/lib/variants
{ echo "# minb0"; python tools/sweep.py --only dasin,dacos --reps 10 --cooldown 1;
  for v in am5 am6 am8 au2; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/sweep.py --only dasin,dacos --reps 10 --cooldown 1; done; echo "# minb0_again"; python tools/sweep.py --only dasin,dacos --reps 10 --cooldown 1; } > gpurun_out/ab_asin64_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_asin64_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:].split(' ')[0]
    elif line.startswith('{'):
        r=json.loads(line); tab.setdefault(r['symbol'],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items(): print("%-14s %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items())))
P
