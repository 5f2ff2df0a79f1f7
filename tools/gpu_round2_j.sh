#!/bin/bash
mkdir -p gpurun_out
V=numericals_b200/lib/variants
{ echo "# default"; python tools/sweep.py --only sasinh,sacosh,satanh,i8hmax,u8hmax,i8dot,i8sum,spow --reps 10;
  for v in hyu4 hym5 hym4u4 hym6; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/sweep.py --only sasinh,sacosh,satanh --reps 10; done; } > gpurun_out/ab_hyp_r02.txt 2>&1
python - <<'P'
import json
for line in open('gpurun_out/ab_hyp_r02.txt'):
    if line.startswith('#'): print(line.strip())
    elif line.startswith('{'):
        r=json.loads(line); print("   %-12s %.4f ms  %.3f of measured" % (r['symbol'], r['ms'], r['frac_of_measured_peak']))
P
timeout 600 python -m pytest tests/test_transcendental_gpu.py tests/test_reduce_gpu.py tests/test_next_rows_gpu.py tests/test_bmas_gpu.py -q 2>&1 | tail -4
