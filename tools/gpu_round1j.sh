#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python tools/opbench.py undelta_pack,undelta_pack_untranspose,transpose_delta_pack,untranspose,transpose,pack > gpurun_out/opbench_fused_r1j.log 2>&1; grep -E "u32|u64" gpurun_out/opbench_fused_r1j.log
