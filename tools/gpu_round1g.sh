#!/bin/bash
# 2-GPU validation of the torchrun path (both workloads) + reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 exit $?"; tail -3 gpurun_out/bench_n2.err; cut -c1-300 gpurun_out/bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload scaling --steps 3 --warmup 3 > gpurun_out/bench_scaling_n2.json 2> gpurun_out/bench_scaling_n2.err; echo "scaling n2 exit $?"; cut -c1-300 gpurun_out/bench_scaling_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 exit $?"; cut -c1-200 gpurun_out/bench_ref_n2.json
