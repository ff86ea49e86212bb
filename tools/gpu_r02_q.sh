#!/bin/bash
# round-2 call Q (2 GPUs): context family + direct path on two real devices (page-locked buffers registered portable + mapped),
# NCCL shard test, N=2 bench as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_ctx.py tests/test_gpu_multi.py tests/test_gpu_host_path.py tests/test_gpu_cpp_traits.py -x -q -m gpu > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_q.log
timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_q.json 2> gpurun_out/bench_n2_q.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n2_q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_q.json').read().strip().splitlines()[-1])
e=d['e2e']; o=d['roofline']['other']
print('value',d['value'],'frac',d['roofline']['frac'])
print('e2e',e['value'],'ceiling',e['link_ceiling']['value'],e['frac_of_ceiling'],'scan',e['scan_filter']['value'])
print('shard',o['sharded_batch_u32_w16']); print('verify',o['sharded_verify'])
PY
