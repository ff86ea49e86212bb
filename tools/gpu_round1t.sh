#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/opbench.py pack,for_pack > gpurun_out/opbench_pack_tma.log 2>&1; cat gpurun_out/opbench_pack_tma.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pack_every_width or for_family" > gpurun_out/sanitizer_racecheck_tma_pack.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck_tma_pack.log; tail -4 gpurun_out/sanitizer_racecheck_tma_pack.log
