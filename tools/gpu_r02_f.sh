#!/bin/bash
# round-2 call F (1 GPU): full suite on the chosen defaults, u16 delta occupancy A/B, compute-sanitizer on the new kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_f.log
echo "== undelta_pack u16 FLB_U16_DELTA_OCC=0"; FLB_U16_DELTA_OCC=0 timeout 300 python tools/opbench.py undelta_pack --types 16 2>&1 | tee gpurun_out/opbench_u16_delta_occ0.txt
echo "== undelta_pack u16 FLB_U16_DELTA_OCC=1"; FLB_U16_DELTA_OCC=1 timeout 300 python tools/opbench.py undelta_pack --types 16 2>&1 | tee gpurun_out/opbench_u16_delta_occ1.txt
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_ctx.py tests/test_gpu_host_path.py -q -m gpu -x > gpurun_out/sanitizer_memcheck_r02.txt 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/sanitizer_memcheck_r02.txt; tail -3 gpurun_out/sanitizer_memcheck_r02.txt
timeout 1500 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity.py -q -m gpu -x -k "select or filter or for_pack_auto or untranspose or transpose_delta or fused" > gpurun_out/sanitizer_racecheck_r02.txt 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/sanitizer_racecheck_r02.txt; tail -3 gpurun_out/sanitizer_racecheck_r02.txt
