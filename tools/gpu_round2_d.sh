#!/bin/bash
mkdir -p gpurun_out
V=numericals_b200/lib/variants
{ echo "# default (both polynomials, Payne-Hanek called)"; python tools/sweep.py --only ssin,scos,stan,stanh --reps 10;
  for v in selcoef tanhm6 tanhm5 trigm6 trigm0; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/sweep.py --only ssin,scos,stan,stanh --reps 10; done; } > gpurun_out/ab_f32_r02b.txt 2>&1
python tools/time_c3.py > gpurun_out/time_c3_r02.txt 2>&1
python - <<'P'
import json
for line in open('gpurun_out/ab_f32_r02b.txt'):
    if line.startswith('#'): print(line.strip())
    elif line.startswith('{'):
        r=json.loads(line); print("   %-12s %.4f ms  %.3f of measured" % (r['symbol'], r['ms'], r['frac_of_measured_peak']))
P
cat gpurun_out/time_c3_r02.txt
