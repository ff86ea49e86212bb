#!/bin/bash
# round-2 call O (1 GPU): direct path for mid-size host calls on page-locked buffers — parity, A/B (tools/midbench.cpp), latency
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host_path.py -x -q -m gpu > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_o.log
{ echo "== default (direct path on)"; timeout 300 build/midbench 200; echo "== FLB_DIRECT_MAX=0 (chunked copy pipeline only)"; FLB_DIRECT_MAX=0 timeout 300 build/midbench 200; echo "== FLB_DIRECT_MAX=1073741824"; FLB_DIRECT_MAX=1073741824 timeout 300 build/midbench 200; } > gpurun_out/midbench_o.txt 2>&1; cat gpurun_out/midbench_o.txt
timeout 120 build/latbench 2000 > gpurun_out/latbench_o.txt 2>&1; tail -8 gpurun_out/latbench_o.txt
