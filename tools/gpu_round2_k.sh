#!/bin/bash
mkdir -p gpurun_out
V=numericals_b200/lib/variants
{ echo "# default"; python tools/sweep.py --only sasinh,sacosh,satanh --reps 10;
  for v in cvt1 cvt2 cvt3; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/sweep.py --only sasinh,sacosh,satanh --reps 10; done; } > gpurun_out/ab_hyp_r02c.txt 2>&1
python - <<'P'
import json
for line in open('gpurun_out/ab_hyp_r02c.txt'):
    if line.startswith('#'): print(line.strip())
    elif line.startswith('{'):
        r=json.loads(line); print("   %-12s %.4f ms  %.3f of measured" % (r['symbol'], r['ms'], r['frac_of_measured_peak']))
P
