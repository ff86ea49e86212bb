#!/bin/bash
# 2-GPU validation on one box: NCCL scatter / scan / gather test (tools/mgpu_scan.py under torchrun) + the thread-safety test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_threads.py -q -m gpu > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_multi.log
