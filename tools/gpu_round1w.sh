#!/bin/bash
# round 1, session 2, call 2: scan kernels after the predicate rewrite; u8 fused original-order chains A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scan.py -q -m gpu -x > gpurun_out/pytest_scan.log 2>&1; echo "scan pytest exit $?"; tail -5 gpurun_out/pytest_scan.log
for v in warp slice; do
  FLB_U8_ORIG=$v timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused_original_order" > gpurun_out/pytest_u8orig_$v.log 2>&1; echo "u8 orig $v pytest exit $?"; tail -3 gpurun_out/pytest_u8orig_$v.log
  FLB_U8_ORIG=$v timeout 600 python tools/opbench.py undelta_pack_untranspose,transpose_delta_pack --types 8 > gpurun_out/opbench_u8orig_$v.log 2>&1; echo "opbench $v exit $?"; cat gpurun_out/opbench_u8orig_$v.log
done
timeout 600 python tools/opbench.py unpack_filter,unpack_select_25pct > gpurun_out/opbench_scan2.log 2>&1; echo "opbench exit $?"; grep -E "filter" gpurun_out/opbench_scan2.log
