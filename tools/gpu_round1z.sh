#!/bin/bash
# round 1, session 2, call 5: in-place compare filter (u32/u64), for_pack_auto
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python tools/opbench.py unpack_filter,for_pack,for_pack_auto --types 32,64 > gpurun_out/opbench_z.log 2>&1; echo "opbench exit $?"; cat gpurun_out/opbench_z.log
