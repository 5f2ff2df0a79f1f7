#!/bin/bash
# run #2 of round 2 (2 GPUs): multi-GPU tests, 2-rank bench, PCIe / host-copy ceilings, pageable-path A/B
mkdir -p gpurun_out
export NUK_COMM_TIMEOUT_MS=10000
nvidia-smi topo -m > gpurun_out/topo2_r02.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu.py tests/test_runtime_gpu.py tests/test_configs_gpu.py::test_c3_broadcast_shapes_bit_exact -q --maxfail=10 > gpurun_out/gputest_r02b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r02b.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02b_n2.json 2> gpurun_out/bench_r02b_n2.err; echo "bench2 rc=$?"
timeout 300 python tools/pciebench_multi.py > gpurun_out/pciebench_multi_r02b.txt 2>&1
for t in 2 4 8 15; do NUK_STAGE_THREADS=$t timeout 120 python tools/time_pageable.py; done > gpurun_out/time_pageable_r02b.txt 2>&1
NUK_STAGE_BOUNCE=0 timeout 120 python tools/time_pageable.py >> gpurun_out/time_pageable_r02b.txt 2>&1
tail -15 gpurun_out/gputest_r02b.log; tail -3 gpurun_out/bench_r02b_n2.err; cat gpurun_out/pciebench_multi_r02b.txt gpurun_out/time_pageable_r02b.txt
