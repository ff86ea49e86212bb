#!/bin/bash
# round-2 call V (8 GPUs): final library — context family + direct path on 8 real devices, bench exactly as the driver launches it at N=8
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_ctx.py tests/test_gpu_multi.py tests/test_gpu_host_path.py -x -q -m gpu > gpurun_out/pytest_gpu_v.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_v.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29715 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/bench_n${NG}_v.json 2> gpurun_out/bench_n${NG}_v.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n${NG}_v.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${NG}_v.json').read().strip().splitlines()[-1])
e=d['e2e']; o=d['roofline']['other']
print('value',d['value'],'frac',d['roofline']['frac'])
print('e2e',e['value'],'ms',e['ms_per_step'],'ceiling',e['link_ceiling']['value'],'frac_of_ceiling',e['frac_of_ceiling'],'scan',e['scan_filter']['value'], e.get('ref_bench_shape',{}).get('decompress_us'))
print('shard',o['sharded_batch_u32_w16']['strong_scaling_efficiency'], o['sharded_batch_u32_w16']['Gints']); print('verify',o['sharded_verify']['match'], o['sharded_verify'].get('oracle_sampled_blocks_match')); print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'],d['cpu_baseline']['scan_filter_Gints'])
PY
