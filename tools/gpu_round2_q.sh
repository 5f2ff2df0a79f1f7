#!/bin/bash
mkdir -p gpurun_out
V=numericals_b200/lib/variants
T=tan,asin,acos,atan,sinh,cosh,asinh,acosh,atanh,pow,atan2
{ echo "# default"; python tools/sweep.py --only $T,dsin,dcos,dtanh,dlog,dexp --reps 10;
  echo "# f32u8"; NUK_LIB=$V/libnuk_f32u8.so python tools/sweep.py --only ^s --reps 10;
  echo "# f32u6"; NUK_LIB=$V/libnuk_f32u6.so python tools/sweep.py --only ^s --reps 10;
  echo "# f64u4"; NUK_LIB=$V/libnuk_f64u4.so python tools/sweep.py --only ^d --reps 10;
  echo "# f64u3"; NUK_LIB=$V/libnuk_f64u3.so python tools/sweep.py --only ^d --reps 10; } > gpurun_out/ab_unroll_all_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_unroll_all_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:]
    elif line.startswith('{'):
        r=json.loads(line)
        if r['kind'] in ('unary','binary') and r['symbol'][5] in 'sd' and not any(t in r['symbol'] for t in ('add','sub','mul','div','max','min','fabs','round','floor','sqrt','copy')):
            tab.setdefault(r['symbol'],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items():
    if len(v)>1:
        best=max(v,key=lambda c:v[c]); print("%-14s %s  %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items()), ("<<< "+best) if best!='default' and v[best]>v['default']*1.03 else ""))
P
