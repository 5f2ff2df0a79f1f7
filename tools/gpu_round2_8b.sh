#!/bin/bash
# 8-GPU run of the final binary: bench at N = 8 and N = 4 (the multi-GPU tests ran on 8 GPUs earlier in the round:
# profiles/gputest_r02_8gpu.log; the communicator code has not changed since)
mkdir -p gpurun_out
export NUK_COMM_TIMEOUT_MS=15000
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r02_n8_final.json 2> gpurun_out/bench_r02_n8.err; echo "bench8 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_r02_n4_final.json 2> gpurun_out/bench_r02_n4.err; echo "bench4 rc=$?"
python - <<'P'
import json
for n in (8,4):
    try:
        d=json.load(open('gpurun_out/bench_r02_n%d_final.json'%n))
        print(n, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],2), 'C5', d['extra']['C5_multiply_sum_2p32_sharded'].get('ms'))
    except Exception as e: print(n, 'ERR', e)
P
