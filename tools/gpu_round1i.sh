#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/opbench.py transpose,untranspose > gpurun_out/opbench_transpose_r1i.log 2>&1; cat gpurun_out/opbench_transpose_r1i.log
# memory checker over the small-size parity tests (all kernels, all types, ragged batches)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log; tail -8 gpurun_out/sanitizer_memcheck.log
