#!/bin/bash
# round 1, session 2, call 3: full GPU test suite (u8 slice default, scan carry fix), filter TMA A/B, u8 op table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
for v in 1 0; do
  FLB_FILTER_TMA=$v timeout 600 python tools/opbench.py unpack_filter --types 16,32,64 > gpurun_out/opbench_filter_tma$v.log 2>&1; echo "filter tma=$v exit $?"; cat gpurun_out/opbench_filter_tma$v.log
done
timeout 600 python tools/opbench.py unpack_filter,undelta_pack_untranspose,transpose_delta_pack --types 8 > gpurun_out/opbench_u8.log 2>&1; cat gpurun_out/opbench_u8.log
