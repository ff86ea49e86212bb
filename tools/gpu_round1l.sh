#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
K=build/kbench
{
for t in 8 16; do
  for op in undelta_pack undelta_packB unpack unpackB; do timeout 100 $K/kb_base $t $op 22 5 1 3; timeout 100 $K/kb_base $t $op 22 5 5 5; done
  timeout 60 $K/kb_base $t delta 22 5; timeout 60 $K/kb_base $t deltaB 22 5; timeout 60 $K/kb_base $t undelta 22 5; timeout 60 $K/kb_base $t undeltaB 22 5
done
} > gpurun_out/kbench_r1l.log 2>&1
grep -v "^#" gpurun_out/kbench_r1l.log
timeout 600 python tools/opbench.py undelta_pack,delta,undelta,undelta_pack_untranspose,transpose_delta_pack,for_pack,unfor_pack 2>&1 | grep -E "u16|u8 " > gpurun_out/opbench_r1l.log; cat gpurun_out/opbench_r1l.log
