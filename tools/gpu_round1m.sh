#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
echo "== warp"; timeout 600 python tools/opbench.py transpose,untranspose 2>&1 | tee gpurun_out/opbench_transpose_warp.log
echo "== tile"; FLB_TRANSPOSE=tile timeout 600 python tools/opbench.py transpose,untranspose 2>&1 | tee gpurun_out/opbench_transpose_tile.log
timeout 300 python tools/opbench.py undelta_pack 2>&1 | grep "u8 " | tee gpurun_out/opbench_u8_delta.log
