#!/bin/bash
# round-2 call T (1 GPU): row-slice select for u8 / u16 — parity (incl. sanitizer) and A/B against the warp-block kernel
mkdir -p gpurun_out
FLB_SELECT_SLICE=1 timeout 600 python -m pytest tests/test_gpu_scan.py -x -q -m gpu -k select > gpurun_out/pytest_gpu_t.log 2>&1; echo "pytest slice exit $?"; tail -3 gpurun_out/pytest_gpu_t.log
CS=/usr/local/cuda/bin/compute-sanitizer
FLB_SELECT_SLICE=1 timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py -q -m gpu -x -k "select and (8 or 16)" > gpurun_out/sanitizer_memcheck_select_slice.txt 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/sanitizer_memcheck_select_slice.txt
FLB_SELECT_SLICE=1 timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py -q -m gpu -x -k "select_every_width and (8 or 16)" > gpurun_out/sanitizer_racecheck_select_slice.txt 2>&1; echo "racecheck exit $?"; tail -3 gpurun_out/sanitizer_racecheck_select_slice.txt
FLB_SELECT_SLICE=0 timeout 400 python tools/select_sweep.py 8 16 > gpurun_out/select_sweep_t_warp.txt 2>&1; echo "warp $?"
FLB_SELECT_SLICE=1 timeout 400 python tools/select_sweep.py 8 16 > gpurun_out/select_sweep_t_slice.txt 2>&1; echo "slice $?"
paste gpurun_out/select_sweep_t_warp.txt gpurun_out/select_sweep_t_slice.txt | awk '{printf "%s %s warp=%s slice=%s  x%.2f  GB/s=%s\n",$1,$2,$3,$7,$3/$7,$8}'
