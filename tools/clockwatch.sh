#!/bin/bash
# tools/clockwatch.sh SYMS -- SM clock / power / throttle reasons sampled every 50 ms while tools/sweep.py times SYMS
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader,nounits -lms 50 > gpurun_out/clockwatch.csv &
W=$!
sleep 1
python tools/sweep.py --only $1 --reps 200 2>&1 | cut -c1-120
kill $W
awk -F', ' '{c[$1]++; if ($3>p) p=$3; if ($4 ~ /Active/ && $4 !~ /Not/) cap++} END {for (k in c) printf "sm_mhz %s: %d samples\n", k, c[k]; printf "max power %.0f W, sw_power_cap active in %d samples of %d\n", p, cap, NR}' gpurun_out/clockwatch.csv | sort -k2 -n
