#!/bin/bash
# quick check of the scan kernels: parity + filter op table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py -q -m gpu > gpurun_out/pytest_scan.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_scan.log
timeout 600 python tools/opbench.py unpack_filter > gpurun_out/opbench_q.log 2>&1; echo "opbench exit $?"; cat gpurun_out/opbench_q.log
