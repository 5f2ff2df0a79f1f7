#!/bin/bash
# tools/profile_round2b.sh -- the final-binary refresh of profiles/*_r02* in one gpurun call (1 GPU): GPU tests, smoke,
# full sweep, region-cost timing, bench (reference arm, own arm), then -- each only after the same command has exited
# 0 without a profiler -- the ncu launch list of the TIMED REGION of bench.py, --set full captures of the C2 kernels and
# of the kernels behind the other BASELINE configs, and the dynamic pipe statistics of the transcendental maps.
set -x
o=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $o/gputest_r02.log 2>&1; echo pytest=$?
python __graft_entry__.py smoke > $o/smoke_r02.log 2>&1; echo smoke=$?
python tools/sweep.py --out $o/sweep_r02.jsonl > $o/sweep_r02.log 2>&1; echo sweep=$?
python tools/sweep.py --cooldown 1 --only ^s,^d --out $o/sweep_r02_cooldown.jsonl > $o/sweep_r02_cooldown.log 2>&1; echo sweepcool=$?
python tools/time_regions.py > $o/time_regions_r02.txt 2>&1; echo regions=$?
python bench.py --impl reference --steps 3 --warmup 1 > $o/bench_ref_r02_n1.json 2> $o/bench_ref_r02.err; echo ref=$?
python bench.py > $o/bench_r02_n1.json 2> $o/bench_r02.err; echo bench=$?
python bench.py --timed-only --steps 3 --warmup 3 > $o/bench_timed.json 2> $o/bench_timed.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/ncu_launches_timed_r02.csv \
      python bench.py --timed-only --steps 3 --warmup 3 > $o/ncu_timed.log 2>&1; echo launches=$?
python tools/prof_c2.py 1 > $o/prof_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_map_flat -o $o/prof_c2_full_r02 -f \
      python tools/prof_c2.py 1 > $o/ncu_full.log 2>&1; echo c2full=$?
python tools/ncu_summary.py $o/prof_c2_full_r02.ncu-rep $o/ncu_c2_full_r02.json --traffic $o/traffic_r02.json > $o/ncu_c2_summary.log 2>&1
ncu -i $o/prof_c2_full_r02.ncu-rep --page details --csv 2>/dev/null | grep -E "Issue Slots Busy|Executed Ipc|Registers Per|Theoretical Occupancy|Achieved Occupancy|DRAM Throughput|Duration|Warp Cycles Per Issued" > $o/ncu_c2_details_r02.csv
rm -f $o/prof_c2_full_r02.ncu-rep   # gpurun_out is capped at 64 MiB: the summaries travel, the report does not
python tools/prof_misc.py > $o/prof_misc_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:k_map_flat2|k_reduce_flat|k_reduce_cols_tma|k_reduce_rows|k_dot" -c 16 -o $o/prof_misc_full_r02 -f \
      python tools/prof_misc.py > $o/ncu_misc.log 2>&1; echo miscfull=$?
python tools/ncu_summary.py $o/prof_misc_full_r02.ncu-rep $o/ncu_misc_full_r02.json --traffic $o/traffic_misc_r02.json > $o/ncu_misc_summary.log 2>&1
rm -f $o/prof_misc_full_r02.ncu-rep
SY=sexp,slog,ssin,scos,stan,sasin,sacos,satan,ssinh,scosh,stanh,sasinh,sacosh,satanh,spow,satan2,dexp,dlog,dsin,dcos,dtan,dasin,dacos,datan,dsinh,dcosh,dtanh,dasinh,dacosh,datanh,dpow,datan2
python tools/prof_ops.py 24 $SY > $o/prof_ops_plain.log 2>&1 && \
  ncu --section ComputeWorkloadAnalysis --section InstructionStats --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats \
      --clock-control none -k regex:k_map_flat -o $o/pipes_all -f python tools/prof_ops.py 24 $SY > $o/ncu_pipes.log 2>&1; echo pipes=$?
ncu -i $o/pipes_all.ncu-rep --page raw --csv > $o/pipes_all.csv 2>/dev/null; python tools/pipes_table.py $o/pipes_all.csv 24 > $o/pipes_all_r02.txt; rm -f $o/pipes_all.ncu-rep $o/pipes_all.csv
tools/_build/c1_latency numericals_b200/lib/libnuk.so > $o/c1_latency_r02.txt 2>&1
python tools/time_c3.py > $o/time_c3_r02_final.txt 2>&1
tail -3 $o/gputest_r02.log; cat $o/smoke_r02.log | tail -1; head -c 600 $o/bench_r02_n1.json
