#!/bin/bash
# first GPU call: parity tests + kernel micro-benchmarks
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu_info.txt; free -g >> gpurun_out/gpu_info.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
K=build/kbench
{
timeout 120 $K/kb_base 32 copy 20 10
timeout 300 $K/kb_base 32 unpack 20 10
for v in st1 st2 ld1 ld3 pf4 pf16 t128 t512; do
  for w in 1 8 16 24 32; do timeout 60 $K/kb_$v 32 unpack 20 10 $w $w; done
done
timeout 120 $K/kb_base 32 pack 20 10
timeout 120 $K/kb_base 32 undelta_pack 20 10 8 8
timeout 300 $K/kb_base 64 unpack 19 5
timeout 120 $K/kb_base 16 unpack 21 5
timeout 120 $K/kb_base 8 unpack 22 5
} > gpurun_out/kbench_r1a.log 2>&1
tail -60 gpurun_out/kbench_r1a.log
