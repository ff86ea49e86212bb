#!/bin/bash
# 1/2/4/8-GPU run of bench.py exactly as the driver launches it (weak sweep + roofline.other incl. the strong-scaling
# shard of configs[4] + e2e with its copy-only ceiling), for as many GPUs as the box has
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_scale.txt
NG=$(nvidia-smi -L | wc -l)
P=29600
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  P=$((P+1))
  if [ $n -eq 1 ]; then
    timeout 800 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-ops > gpurun_out/scale_sweep_n$n.json 2> gpurun_out/scale_sweep_n$n.err
  else
    timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_sweep_n$n.json 2> gpurun_out/scale_sweep_n$n.err
  fi
  echo "n=$n: $(python -c "
import json
d=json.loads(open('gpurun_out/scale_sweep_n$n.json').read().strip().splitlines()[-1])
e=d['e2e']; o=d['roofline']['other']['sharded_batch_u32_w16']
print(d['value'],d['unit'],'| e2e',e['value'],'ceiling',e['link_ceiling']['value'],'frac',e['frac_of_ceiling'],'nodes',e['device_numa_node'],e['buffer_numa_node'],'| shard ms',o['ms'],'eff',o['strong_scaling_efficiency'])" 2>&1)"
done
