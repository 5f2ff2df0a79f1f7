#!/bin/bash
mkdir -p gpurun_out
V=numericals_b200/lib/variants
T=sin,cos,tan,asin,acos,atan,sinh,cosh,tanh,asinh,acosh,atanh,exp,log,pow,atan2
{ echo "# default"; python tools/sweep.py --only $T --reps 10;
  echo "# ws64"; NUK_LIB=$V/libnuk_ws64.so python tools/sweep.py --only ^d --reps 10;
  echo "# ws32"; NUK_LIB=$V/libnuk_ws32.so python tools/sweep.py --only ^s --reps 10; } > gpurun_out/ab_ws_r02.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/ab_ws_r02.txt'):
    if line.startswith('#'): cur=line.strip()[2:]
    elif line.startswith('{'):
        r=json.loads(line)
        if r['kind'] in ('unary','binary') and r['symbol'][5] in 'sd' and not any(t in r['symbol'] for t in ('add','sub','mul','div','max','min','fabs','round','floor','sqrt','copy')):
            tab.setdefault(r['symbol'],{})[cur if cur=='default' else 'warp']=r['frac_of_measured_peak']
for k,v in tab.items():
    if len(v)>1: print("%-14s default %.3f  warp-split %.3f  %s" % (k, v['default'], v['warp'], "<<<" if v['warp']>v['default']*1.03 else ""))
P
