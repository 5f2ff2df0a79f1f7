#!/bin/bash
# kRegionSplit (one vote per tile: polynomial for warps entirely inside the region, the old logarithm-path pack split
# for every other warp): tests, then the region-cost timing for packs-per-thread variants
mkdir -p gpurun_out
V=numericals_b200/lib/variants
timeout 900 python -m pytest tests/test_transcendental_gpu.py tests/test_table_kernels_gpu.py -q -x > gpurun_out/gputest_trans_r02f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_trans_r02f.log; tail -3 gpurun_out/gputest_trans_r02f.log
{ echo "# default"; python tools/time_regions.py;
  for v in as2 as8 ih2 ihm0; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/time_regions.py; done; } > gpurun_out/time_regions_r02_ab.txt 2>&1
python - <<'P'
import json
cur=None; tab={}
for line in open('gpurun_out/time_regions_r02_ab.txt'):
    if line.startswith('#'): cur=line.strip()[2:].split(' ')[0]
    elif line.startswith('{'):
        r=json.loads(line); tab.setdefault(r['symbol']+' '+r['arguments'].split(' ')[0],{})[cur]=r['frac_of_measured_peak']
for k,v in tab.items(): print("%-22s %s" % (k, "  ".join("%s %.3f" % kv for kv in v.items())))
P
