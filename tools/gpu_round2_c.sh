#!/bin/bash
# run #3: A/B of the single-float tanh / trig variants, full sweep, C1 from C, rows-kernel shapes
mkdir -p gpurun_out
V=numericals_b200/lib/variants
{ echo "# default"; python tools/sweep.py --only ssin,scos,stan,stanh,sexp,slog --reps 10;
  for v in tanhfast tanhfast8 tanh6 trigfast trigfast0; do echo "# $v"; NUK_LIB=$V/libnuk_$v.so python tools/sweep.py --only ssin,scos,stan,stanh --reps 10; done; } > gpurun_out/ab_f32_r02.txt 2>&1
python tools/sweep.py --reps 10 --out gpurun_out/sweep_r02a.jsonl > gpurun_out/sweep_r02a.log 2>&1
tools/_build/c1_latency numericals_b200/lib/libnuk.so > gpurun_out/c1_latency_r02.txt 2>&1
timeout 600 python bench.py --no-cpu --e2e-steps 0 --steps 10 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err
grep -v "^{" gpurun_out/ab_f32_r02.txt | head -3; python - <<'P'
import json
cur=None
for line in open('gpurun_out/ab_f32_r02.txt'):
    if line.startswith('#'): cur=line.strip(); print(cur)
    elif line.startswith('{'):
        r=json.loads(line); print("   %-12s %.4f ms  %.3f of measured" % (r['symbol'], r['ms'], r['frac_of_measured_peak']))
P
cat gpurun_out/c1_latency_r02.txt
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r02c.json'))
for k in ('C3a_col_plus_row_8192','C3b_matrix_times_row_4096','C3b_matrix_times_row_16384'):
    print(k, d['extra'][k])
print({k:(round(v['ms'],4), round(v['frac_of_measured_peak'],3)) for k,v in d['per_ufunc'].items()})
P
