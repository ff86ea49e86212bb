#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/opbench.py block_minmax 2>&1 | tee gpurun_out/opbench_minmax.log
timeout 600 python tools/refbench.py 2>&1 | tee gpurun_out/refbench.log
