#!/bin/bash
# round-2 call L (1 GPU): full-width sweep of the select variants on one box (choice of blocks per warp per type / width)
mkdir -p gpurun_out
for mode in "FLB_SELECT=warp2" "FLB_SELECT=warp2 FLB_SELECT_NB=4" "FLB_SELECT=warp2 FLB_SELECT_NB=8"; do
  env $mode timeout 600 python -m pytest tests/test_gpu_scan.py -x -q -m gpu -k select > gpurun_out/pytest_gpu_l_sel.log 2>&1; echo "pytest [$mode] exit $?"; tail -1 gpurun_out/pytest_gpu_l_sel.log
done
timeout 400 python tools/select_sweep.py > gpurun_out/select_sweep_default.txt 2>&1; echo "default $?"
FLB_SELECT=warp2 timeout 400 python tools/select_sweep.py > gpurun_out/select_sweep_nb1.txt 2>&1; echo "nb1 $?"
FLB_SELECT=warp2 FLB_SELECT_NB=4 timeout 400 python tools/select_sweep.py > gpurun_out/select_sweep_nb4.txt 2>&1; echo "nb4 $?"
FLB_SELECT=warp2 FLB_SELECT_NB=8 timeout 400 python tools/select_sweep.py > gpurun_out/select_sweep_nb8.txt 2>&1; echo "nb8 $?"
paste gpurun_out/select_sweep_default.txt gpurun_out/select_sweep_nb1.txt gpurun_out/select_sweep_nb4.txt gpurun_out/select_sweep_nb8.txt | awk '{print $1,$2,$3,$7,$11,$15}'
