#!/bin/bash
# pageable host operands: chunk size x cached / streaming stores into the ring (does a ring that fits the L3 save the
# DRAM round trip?), then the double-float sweep of the adopted settings
mkdir -p gpurun_out
{ for mb in 32 8 4 2 1; do for nt in 1 0; do echo -n "stage_mb=$mb stage_nt=$nt  "; NUK_STAGE_MB=$mb NUK_STAGE_NT=$nt python tools/time_pageable.py; done; done;
  for sl in 4 6; do echo -n "stage_mb=4 stage_nt=0 slots=$sl  "; NUK_STAGE_MB=4 NUK_STAGE_NT=0 NUK_STAGE_SLOTS=$sl python tools/time_pageable.py; done; } > gpurun_out/time_pageable_r02_b.txt 2>&1
cat gpurun_out/time_pageable_r02_b.txt
lscpu | grep -E "Model name|L3|L2|Socket|^CPU\(s\)" >> gpurun_out/time_pageable_r02_b.txt
python tools/sweep.py --only dsin,dcos,dtan,dasin,dacos,datan,datan2,dsinh,dcosh,dtanh,dasinh,dacosh,datanh,dpow --reps 10 > gpurun_out/sweep_w.jsonl 2>&1
python - <<'P'
import json
for l in open('gpurun_out/sweep_w.jsonl'):
    if l.startswith('{'):
        r=json.loads(l); print("%-14s %.3f ms %.3f" % (r['symbol'], r['ms'], r['frac_of_measured_peak']))
P
