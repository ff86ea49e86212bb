#!/bin/bash
# 1/2/4/8-GPU scaling of bench.py (weak: configs[1]) and of the sharded workload (strong: configs[4])
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_scale.txt
P=29600
for n in 1 2 4 8; do
  P=$((P+1))
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_sweep_n$n.json 2> gpurun_out/scale_sweep_n$n.err
    timeout 300 python bench.py --gpus 1 --workload scaling --steps 3 --warmup 3 > gpurun_out/scale_shard_n$n.json 2> gpurun_out/scale_shard_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_sweep_n$n.json 2> gpurun_out/scale_sweep_n$n.err
    P=$((P+1))
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --workload scaling --steps 3 --warmup 3 > gpurun_out/scale_shard_n$n.json 2> gpurun_out/scale_shard_n$n.err
  fi
  echo "n=$n sweep: $(python -c "import json;d=json.loads(open('gpurun_out/scale_sweep_n$n.json').read().strip().splitlines()[-1]);print(d['value'],d['unit'],d['gbps'],'GB/s e2e',d['e2e']['value'] if d['e2e'] else None)" 2>&1)"
  echo "n=$n shard: $(python -c "import json;d=json.loads(open('gpurun_out/scale_shard_n$n.json').read().strip().splitlines()[-1]);print(d['value'],d['unit'],d['gbps'],'GB/s',d['ms_per_step'],'ms')" 2>&1)"
done
