#!/bin/bash
# 8-GPU run: topology, multi-GPU tests (8 devices in one process; 8 processes over cudaIpc), PCIe ceiling by GPU count,
# bench at N = 8 and N = 4, the reference arm on this box's cores
mkdir -p gpurun_out
export NUK_COMM_TIMEOUT_MS=15000
{ nvidia-smi topo -m; nproc; free -g | head -2; (numactl -H 2>/dev/null || ls /sys/devices/system/node | grep node);
  for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; } > gpurun_out/topo8_r02.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu.py tests/test_runtime_gpu.py::test_thread_moving_between_devices_keeps_its_scratch_apart -q > gpurun_out/gputest_r02_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r02_8gpu.log
timeout 300 python tools/pciebench_multi.py > gpurun_out/pciebench_multi_r02_8gpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r02_n8.json 2> gpurun_out/bench_r02_n8.err; echo "bench8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_r02_n4.json 2> gpurun_out/bench_r02_n4.err; echo "bench4 rc=$?"
timeout 300 python bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_r02_8gpubox.json 2>> gpurun_out/bench_r02_n8.err
tail -5 gpurun_out/gputest_r02_8gpu.log; cat gpurun_out/pciebench_multi_r02_8gpu.txt; tail -3 gpurun_out/bench_r02_n8.err
python - <<'P'
import json
for n in (8,4):
    try:
        d=json.load(open('gpurun_out/bench_r02_n%d.json'%n))
        print(n, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['pcie_gbs_each_way_per_gpu'],1), 'pageable', round(d['e2e']['pageable']['value'],2), d.get('host_affinity'))
        print('  C5', {k:d['extra']['C5_multiply_sum_2p32_sharded'].get(k) for k in ('ms','algorithmic_gbs_aggregate','sum_only_ms')})
        print('  C4', {k:round(v['ms'],4) for k,v in d['extra']['C4_sharded_f64_16384sq'].items() if isinstance(v,dict)})
    except Exception as e: print(n, 'ERR', e)
print(open('gpurun_out/bench_ref_r02_8gpubox.json').read()[:300])
P
