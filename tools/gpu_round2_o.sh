#!/bin/bash
mkdir -p gpurun_out
V=numericals_b200/lib/variants
{ for rep in 1 2 3; do
  echo "# default (numerator = D - 2e) run $rep"; python tools/sweep.py --only stanh,stan --reps 10;
  echo "# oldnum run $rep"; NUK_LIB=$V/libnuk_oldnum.so python tools/sweep.py --only stanh,stan --reps 10; done; } > gpurun_out/ab_tanhnum_r02.txt 2>&1
python - <<'P'
import json
for line in open('gpurun_out/ab_tanhnum_r02.txt'):
    if line.startswith('#'): print(line.strip())
    elif line.startswith('{'):
        r=json.loads(line)
        if r['symbol'] in ('BMAS_stanh','BMAS_stan'): print("   %-12s %.4f ms  %.3f of measured" % (r['symbol'], r['ms'], r['frac_of_measured_peak']))
P
timeout 600 python -m pytest tests/test_transcendental_gpu.py -q 2>&1 | tail -2
