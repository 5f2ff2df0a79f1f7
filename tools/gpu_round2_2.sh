#!/bin/bash
# 2-GPU run of the final binary: the whole GPU suite (multi-GPU tests included: 2 devices in one process, 2 processes
# over cudaIpc, the absent-peer timeout), then bench at N = 2
mkdir -p gpurun_out
export NUK_COMM_TIMEOUT_MS=15000
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gputest_r02_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r02_2gpu.log
tail -5 gpurun_out/gputest_r02_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "bench2 rc=$?"
tail -3 gpurun_out/bench_r02_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r02_n2.json'))
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['pcie_gbs_each_way_per_gpu'],1), 'pageable', round(d['e2e']['pageable']['value'],2), d.get('host_affinity'))
print('  C5', {k:d['extra']['C5_multiply_sum_2p32_sharded'].get(k) for k in ('ms','algorithmic_gbs_aggregate','sum_only_ms')})
print('  C4', {k:round(v['ms'],4) for k,v in d['extra']['C4_sharded_f64_16384sq'].items() if isinstance(v,dict)})
P
