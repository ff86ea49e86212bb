#!/bin/bash
# round-2 call R (1 GPU): delta filter with the lo subtraction folded into the carry (parity + timing), compute-sanitizer over
# the final library (scan, host path incl. the direct path, context family)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scan.py -x -q -m gpu > gpurun_out/pytest_gpu_r.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_r.log
timeout 300 python tools/opbench.py undelta_pack_filter 2>&1 | tee gpurun_out/opbench_dfilter_r.txt
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py tests/test_gpu_host_path.py tests/test_gpu_ctx.py tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -x > gpurun_out/sanitizer_memcheck_final.txt 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/sanitizer_memcheck_final.txt; tail -4 gpurun_out/sanitizer_memcheck_final.txt
timeout 1500 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py -q -m gpu -x -k "filter or select or pipeline" > gpurun_out/sanitizer_racecheck_final.txt 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/sanitizer_racecheck_final.txt; tail -4 gpurun_out/sanitizer_racecheck_final.txt
