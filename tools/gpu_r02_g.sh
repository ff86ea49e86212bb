#!/bin/bash
# round-2 call G (8 GPUs): placement probe, context family on 8 real devices, bench exactly as the driver launches it at N=8
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
{ nvidia-smi topo -m; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; timeout 500 build/numa_probe 28; } > gpurun_out/numa_probe_n$NG.txt 2>&1
grep -E "Cpus_allowed|Mems_allowed|node[0-9] cpus|dev[0-9] pci|all on|spread|first " gpurun_out/numa_probe_n$NG.txt
timeout 900 python -m pytest tests/test_gpu_ctx.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu_g.log
timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/bench_n${NG}_g.json 2> gpurun_out/bench_n${NG}_g.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n${NG}_g.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${NG}_g.json').read().strip().splitlines()[-1])
e=d['e2e']; o=d['roofline']['other']
print('value',d['value'],'frac',d['roofline']['frac'])
print('e2e',e['value'],'ms',e['ms_per_step'],'ceiling',e['link_ceiling']['value'],'frac_of_ceiling',e['frac_of_ceiling'],'nodes',e['device_numa_node'],e['buffer_numa_node'],'scan',e['scan_filter'])
print('shard',o['sharded_batch_u32_w16']); print('verify',o['sharded_verify']); print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'],d['cpu_baseline']['scan_filter_Gints'])
PY
