#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "transpose or fused or delta_family" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log; tail -8 gpurun_out/sanitizer_racecheck.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log; tail -6 gpurun_out/sanitizer_memcheck.log
