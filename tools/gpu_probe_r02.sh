#!/bin/bash
# round-2 placement probe (2 or 8 GPUs): what the box exposes + copy-only ceilings per placement
mkdir -p gpurun_out
{
nvidia-smi topo -m
lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"
which numactl && numactl -H
timeout 600 build/numa_probe 29
} > gpurun_out/numa_probe_n$(nvidia-smi -L | wc -l).txt 2>&1
tail -60 gpurun_out/numa_probe_n*.txt
