#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/opbench.py undelta_pack --types 64 2>&1 | tee gpurun_out/opbench_q.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "delta_family" 2>&1 | tail -2
