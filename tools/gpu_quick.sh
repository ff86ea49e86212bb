#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_scan.py -q -m gpu > gpurun_out/pytest_scan.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_scan.log
timeout 600 python tools/opbench.py undelta_pack_filter,undelta_pack > gpurun_out/opbench_q.log 2>&1; echo "opbench exit $?"; cat gpurun_out/opbench_q.log
