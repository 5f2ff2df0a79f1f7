#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_transcendental_gpu.py tests/test_table_kernels_gpu.py tests/test_nd_gpu.py -q -x > gpurun_out/gputest_trans_r02h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_trans_r02h.log; tail -3 gpurun_out/gputest_trans_r02h.log
python tools/time_regions.py 2>&1 | cut -c1-140
