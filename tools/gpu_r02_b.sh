#!/bin/bash
# round-2 call B (2+ GPUs): placement probe, ctx tests on real devices, bench at N=2 (e2e + ceiling + strong-scaling shard)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
{ nvidia-smi topo -m; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; timeout 400 build/numa_probe 29; } > gpurun_out/numa_probe_n$NG.txt 2>&1
tail -45 gpurun_out/numa_probe_n$NG.txt
timeout 900 python -m pytest tests/test_gpu_ctx.py tests/test_gpu_multi.py tests/test_gpu_host_path.py -x -q -m gpu > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu_b.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n2_b.json 2> gpurun_out/bench_n2_b.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n2_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_b.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',json.dumps(d['e2e'])[:1500])
print(json.dumps(d['roofline']['other'])[:3000])
PY
