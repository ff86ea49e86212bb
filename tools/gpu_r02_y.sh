#!/bin/bash
# round-2 call Y (1 GPU): u64 original-order encode chain with full 16-byte shared loads in the gather — parity, then A/B
# against the previous library on the same box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/pytest_gpu_y.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_y.log
echo "== new"; timeout 300 python tools/opbench.py transpose_delta_pack --types 64 2>&1 | tee gpurun_out/opbench_tdp_u64_new.txt
cp fastlanes_b200/lib/libfastlanes_b200.so /tmp/new.so; cp build/lib_old/libfastlanes_b200.so fastlanes_b200/lib/libfastlanes_b200.so
echo "== old"; timeout 300 python tools/opbench.py transpose_delta_pack --types 64 2>&1 | tee gpurun_out/opbench_tdp_u64_old.txt
cp /tmp/new.so fastlanes_b200/lib/libfastlanes_b200.so
echo "== new again"; timeout 300 python tools/opbench.py transpose_delta_pack --types 64 2>&1 | tee gpurun_out/opbench_tdp_u64_new2.txt
