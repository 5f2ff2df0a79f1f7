#!/bin/bash
# double tan: register cap at 4 packs per thread, same box
mkdir -p gpurun_out
V=numericals_b200/lib/variants
{ echo "# minb5"; python tools/sweep.py --only dtan --reps 10 --cooldown 1;
  for v in tm3 tm4 tm6 tm0; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/sweep.py --only dtan --reps 10 --cooldown 1; done; echo "# minb5_again"; python tools/sweep.py --only dtan --reps 10 --cooldown 1; } > gpurun_out/ab_tan64_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_tan64_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:].split(' ')[0]
    elif line.startswith('{'):
        r=json.loads(line); tab.setdefault(r['symbol'],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items(): print("%-14s %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items())))
P
