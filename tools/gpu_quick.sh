#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/opbench.py unpack_select_25pct > gpurun_out/opbench_q.log 2>&1; echo "opbench exit $?"; cat gpurun_out/opbench_q.log
