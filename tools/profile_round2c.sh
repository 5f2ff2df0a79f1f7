#!/bin/bash
# light refresh after a change that does not touch the bench's kernels: GPU tests, smoke, both sweeps, region costs,
# pipe statistics
set -x
o=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $o/gputest_r02.log 2>&1; echo pytest=$?
python __graft_entry__.py smoke > $o/smoke_r02.log 2>&1; echo smoke=$?
python tools/sweep.py --out $o/sweep_r02.jsonl > $o/sweep_r02.log 2>&1; echo sweep=$?
python tools/sweep.py --cooldown 1 --only ^s,^d --out $o/sweep_r02_cooldown.jsonl > $o/sweep_r02_cooldown.log 2>&1; echo sweepcool=$?
python tools/time_regions.py > $o/time_regions_r02.txt 2>&1; echo regions=$?
SY=sexp,slog,ssin,scos,stan,sasin,sacos,satan,ssinh,scosh,stanh,sasinh,sacosh,satanh,spow,satan2,dexp,dlog,dsin,dcos,dtan,dasin,dacos,datan,dsinh,dcosh,dtanh,dasinh,dacosh,datanh,dpow,datan2
python tools/prof_ops.py 24 $SY > $o/prof_ops_plain.log 2>&1 && \
  ncu --section ComputeWorkloadAnalysis --section InstructionStats --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats \
      --clock-control none -k regex:k_map_flat -o $o/pipes_all -f python tools/prof_ops.py 24 $SY > $o/ncu_pipes.log 2>&1; echo pipes=$?
ncu -i $o/pipes_all.ncu-rep --page raw --csv > $o/pipes_all.csv 2>/dev/null; python tools/pipes_table.py $o/pipes_all.csv 24 > $o/pipes_all_r02.txt; rm -f $o/pipes_all.ncu-rep $o/pipes_all.csv
tail -3 $o/gputest_r02.log; tail -1 $o/smoke_r02.log
