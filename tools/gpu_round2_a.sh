#!/bin/bash
# run #1 of round 2: topology probe, GPU suite, bench, launch list of the timed region
mkdir -p gpurun_out
{ nvidia-smi topo -m; echo; lscpu | head -40; echo; (numactl -H 2>/dev/null || ls /sys/devices/system/node); echo; free -g; echo; nproc;
  for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; } > gpurun_out/topo_r02.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/gputest_r02a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r02a.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02a.json 2>> gpurun_out/bench_r02a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_launches_timed_r02a.csv python bench.py --timed-only --steps 3 --warmup 3 > gpurun_out/ncu_timed.log 2>&1
tail -5 gpurun_out/gputest_r02a.log; head -c 1500 gpurun_out/bench_r02a.json
