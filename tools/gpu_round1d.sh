#!/bin/bash
# layout B (warp-block) correctness + speed
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
K=build/kbench
{
timeout 300 $K/kb_u32 32 unpackB 20 10
timeout 120 $K/kb_u32 32 unpack 20 10 16 16
timeout 120 $K/kb_u32 32 undelta_packB 20 10 8 8
timeout 120 $K/kb_u32 32 undelta_pack 20 10 8 8
timeout 120 $K/kb_u32 32 unfor_packB 20 10 8 8
timeout 120 $K/kb_u32 32 pack 20 10
} > gpurun_out/kbench_r1d.log 2>&1
cat gpurun_out/kbench_r1d.log | grep -v "^#"
