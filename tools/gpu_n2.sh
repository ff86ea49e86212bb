#!/bin/bash
# 2-GPU: effect of NUMA-local pinned buffers on the host-buffer (e2e) path, torchrun launch line of the driver
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
for numa in 1 0; do
  FLB_NUMA=$numa timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2971$numa bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n2_numa$numa.json 2> gpurun_out/bench_n2_numa$numa.err; echo "bench n2 numa=$numa exit $?"
done
FLB_NUMA=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n1_numa1.json 2>/dev/null
FLB_NUMA=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n1_numa0.json 2>/dev/null
python - <<'PY'
import json
for f in ("bench_n2_numa1", "bench_n2_numa0", "bench_n1_numa1", "bench_n1_numa0"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    e = d["e2e"]
    print(f, "value", d["value"], "e2e", e["value"], "ms", e["ms_per_step"], "node", e.get("pinned_numa_node"), "scan", e["scan_filter"]["value"])
PY
