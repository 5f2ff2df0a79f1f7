#!/bin/bash
# packs per thread / register caps of the double-float inverse-trig kernels, A/B inside one call
mkdir -p gpurun_out
V=numericals_b200/lib/variants
T=dsin,dcos,dtan,dasin,dacos,datan,datan2
{ echo "# default"; python tools/sweep.py --only $T --reps 10;
  for v in u4 u3 u4m0 a2u2 a2u2m4 a2u2m0; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/sweep.py --only $T --reps 10; done; } > gpurun_out/ab_f64c_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_f64c_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:].split(' ')[0]
    elif line.startswith('{'):
        r=json.loads(line); tab.setdefault(r['symbol'],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items(): print("%-14s %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items())))
P
