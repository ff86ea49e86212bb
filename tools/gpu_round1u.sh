#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/opbench.py delta,undelta,undelta_pack_untranspose > gpurun_out/opbench_delta_tma.log 2>&1; grep -vE "u8 " gpurun_out/opbench_delta_tma.log
