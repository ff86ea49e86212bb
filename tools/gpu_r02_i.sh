#!/bin/bash
# round-2 call I (1 GPU): select third step (select_warp2_kernel) — parity in every mode, then A/B at 25 % selectivity
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest default exit $?"; tail -3 gpurun_out/pytest_gpu_i.log
for mode in "FLB_SELECT=warp2" "FLB_SELECT=warp2 FLB_SELECT_NB=2" "FLB_SELECT=warp2 FLB_SELECT_TMA=1"; do
  env $mode timeout 600 python -m pytest tests/test_gpu_scan.py -x -q -m gpu -k select > gpurun_out/pytest_gpu_i_sel.log 2>&1; echo "pytest [$mode] exit $?"; tail -2 gpurun_out/pytest_gpu_i_sel.log
done
echo "== default"; timeout 300 python tools/opbench.py unpack_select_25pct 2>&1 | tee gpurun_out/opbench_select_i_default.txt
echo "== warp2 nb1"; FLB_SELECT=warp2 timeout 300 python tools/opbench.py unpack_select_25pct 2>&1 | tee gpurun_out/opbench_select_i_warp2.txt
echo "== warp2 nb2"; FLB_SELECT=warp2 FLB_SELECT_NB=2 timeout 300 python tools/opbench.py unpack_select_25pct 2>&1 | tee gpurun_out/opbench_select_i_warp2_nb2.txt
echo "== warp2 tma"; FLB_SELECT=warp2 FLB_SELECT_TMA=1 timeout 300 python tools/opbench.py unpack_select_25pct --types 16,32 2>&1 | tee gpurun_out/opbench_select_i_warp2_tma.txt
