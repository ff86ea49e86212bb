#!/bin/bash
# compute-sanitizer over the session-2 kernels (scan, for_pack_auto, u8 row-slice chains) + u8 minmax group-size A/B
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -x > gpurun_out/sanitizer_memcheck_s2.txt 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/sanitizer_memcheck_s2.txt; tail -4 gpurun_out/sanitizer_memcheck_s2.txt
timeout 1200 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py -q -m gpu -x -k "filter_every_width or select_every_width or pipeline or for_pack_auto" > gpurun_out/sanitizer_racecheck_s2.txt 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/sanitizer_racecheck_s2.txt; tail -4 gpurun_out/sanitizer_racecheck_s2.txt
for g in 8 16 32; do echo "FLB_MINMAX_G8=$g"; FLB_MINMAX_G8=$g timeout 300 python tools/opbench.py block_minmax --types 8; done 2>&1 | tee gpurun_out/opbench_minmax_u8.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k minmax 2>&1 | tail -2
