#!/bin/bash
mkdir -p gpurun_out
V=numericals_b200/lib/variants
{ echo "# default (warp split: trig cap 32 regs, tanh uncapped)"; python tools/sweep.py --only ssin,scos,stan,stanh --reps 10;
  for v in trig0 trig2m6 trig2m0 tanh0 tanh2m8; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/sweep.py --only ssin,scos,stan,stanh --reps 10; done;
  echo "# default again"; python tools/sweep.py --only ssin,scos,stan,stanh --reps 10; } > gpurun_out/ab_warp_r02.txt 2>&1
python - <<'P'
import json
for line in open('gpurun_out/ab_warp_r02.txt'):
    if line.startswith('#'): print(line.strip())
    elif line.startswith('{'):
        r=json.loads(line)
        if r['symbol'] in ('BMAS_ssin','BMAS_scos','BMAS_stan','BMAS_stanh'): print("   %-12s %.4f ms  %.3f of measured" % (r['symbol'], r['ms'], r['frac_of_measured_peak']))
P
timeout 600 python -m pytest tests/test_transcendental_gpu.py tests/test_table_kernels_gpu.py -q 2>&1 | tail -3
