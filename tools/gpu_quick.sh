#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/pytest_q.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_q.log
timeout 600 python tools/opbench.py unpack_cwida,pack_cwida,unpack,pack > gpurun_out/opbench_q.log 2>&1; echo "opbench exit $?"; cat gpurun_out/opbench_q.log
