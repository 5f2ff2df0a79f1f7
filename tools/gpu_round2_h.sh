#!/bin/bash
mkdir -p gpurun_out
V=numericals_b200/lib/variants
T=sin,cos,tan,asin,acos,atan,sinh,cosh,tanh,asinh,acosh,atanh,exp,log,pow,atan2
{ echo "# default"; python tools/sweep.py --only $T --reps 10;
  echo "# rare64"; NUK_LIB=$V/libnuk_rare64.so python tools/sweep.py --only ^d --reps 10;
  echo "# rare32"; NUK_LIB=$V/libnuk_rare32.so python tools/sweep.py --only ^s --reps 10; } > gpurun_out/ab_rare_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_rare_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:]
    elif line.startswith('{'):
        r=json.loads(line)
        if r['kind'] in ('unary','binary') and any(t in r['symbol'] for t in "sin cos tan exp log pow".split()) or 'h' in r['symbol'][5:]:
            tab.setdefault(r['symbol'],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items():
    if len(v)>1: print("%-14s %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items())))
P
