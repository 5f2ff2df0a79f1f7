#!/bin/bash
# integer |x| vs fabs in the double inverse-trig main paths, same box
mkdir -p gpurun_out
V=numericals_b200/lib/variants
T=dasin,dacos,datan,datan2,dsinh,dcosh
{ for rep in 1 2; do echo "# default$rep"; python tools/sweep.py --only $T --reps 10 --cooldown 1; echo "# absb$rep"; NUK_LIB=$V/libnuk_absb.so python tools/sweep.py --only $T --reps 10 --cooldown 1; done; } > gpurun_out/ab_abs_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_abs_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:].split(' ')[0]
    elif line.startswith('{'):
        r=json.loads(line); tab.setdefault(r['symbol'],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items(): print("%-14s %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items())))
P
