#!/bin/bash
mkdir -p gpurun_out
K=build/kbench
{ timeout 300 $K/kb_u32 32 unpackT 20 10; timeout 300 $K/kb_u32 32 unpackB 20 10; } > gpurun_out/kbench_tma.log 2>&1
grep -v "^#" gpurun_out/kbench_tma.log | awk '{print $1,$3,$5,$7}' | paste - - - - | head -40
grep -c MISMATCH gpurun_out/kbench_tma.log
