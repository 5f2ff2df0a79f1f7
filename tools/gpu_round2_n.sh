#!/bin/bash
mkdir -p gpurun_out
V=numericals_b200/lib/variants
{ echo "# default (vote per element)"; python tools/sweep.py --only ssin,scos,stan,stanh --reps 10;
  echo "# wpack (vote per pack)"; NUK_LIB=$V/libnuk_wpack.so python tools/sweep.py --only ssin,scos,stan,stanh --reps 10;
  echo "# default again"; python tools/sweep.py --only ssin,scos,stan,stanh --reps 10;
  echo "# wpack again"; NUK_LIB=$V/libnuk_wpack.so python tools/sweep.py --only ssin,scos,stan,stanh --reps 10; } > gpurun_out/ab_wpack_r02.txt 2>&1
python - <<'P'
import json
for line in open('gpurun_out/ab_wpack_r02.txt'):
    if line.startswith('#'): print(line.strip())
    elif line.startswith('{'):
        r=json.loads(line)
        if r['symbol'] in ('BMAS_ssin','BMAS_scos','BMAS_stan','BMAS_stanh'): print("   %-12s %.4f ms  %.3f of measured" % (r['symbol'], r['ms'], r['frac_of_measured_peak']))
P
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python -m pytest tests/test_nd_gpu.py tests/test_runtime_gpu.py -q 2>&1 | tail -2
