#!/bin/bash
# sinh / cosh without the e^-a guard on the main path, integer |x| in the double inverse-trig main paths: tests + sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_transcendental_gpu.py tests/test_table_kernels_gpu.py -q -x > gpurun_out/gputest_trans_r02g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_trans_r02g.log; tail -3 gpurun_out/gputest_trans_r02g.log
T=dsin,dcos,dtan,dasin,dacos,datan,datan2,dsinh,dcosh,dtanh,dasinh,dacosh,datanh,dpow,dexp,dlog
python tools/sweep.py --only $T --reps 10 > gpurun_out/sweep_y.jsonl 2>&1
python tools/sweep.py --only $T --reps 10 --cooldown 1 > gpurun_out/sweep_y_cool.jsonl 2>&1
python - <<'P'
import json
a={}; 
for f,k in (('gpurun_out/sweep_y.jsonl','b2b'),('gpurun_out/sweep_y_cool.jsonl','cool')):
    for l in open(f):
        if l.startswith('{'):
            r=json.loads(l); a.setdefault(r['symbol'],{})[k]=r['frac_of_measured_peak']
for s,v in a.items(): print("%-14s %s" % (s, "  ".join("%s %.3f" % kv for kv in v.items())))
P
