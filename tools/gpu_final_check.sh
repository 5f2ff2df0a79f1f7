#!/bin/bash
# end-of-round check of the committed tree: GPU tests, smoke, bench (both arms)
o=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $o/gputest_r02.log 2>&1; echo pytest=$?
python __graft_entry__.py smoke > $o/smoke_r02.log 2>&1; echo smoke=$?
python bench.py --impl reference --steps 3 --warmup 1 > $o/bench_ref_r02_n1.json 2> $o/bench_ref_r02.err; echo ref=$?
python bench.py > $o/bench_r02_n1.json 2> $o/bench_r02.err; echo bench=$?
tail -2 $o/gputest_r02.log; tail -1 $o/smoke_r02.log; head -c 400 $o/bench_r02_n1.json
