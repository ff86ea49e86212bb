#!/bin/bash
# round-2 call A (1 GPU): placement probe + full -m gpu suite on the new host path / ctx / select kernel + select opbench
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; timeout 300 build/numa_probe 29; } > gpurun_out/numa_probe_n1.txt 2>&1
tail -40 gpurun_out/numa_probe_n1.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_a.log
timeout 600 python tools/opbench.py unpack_select_25pct,unpack_filter > gpurun_out/opbench_select_a.txt 2>&1; tail -45 gpurun_out/opbench_select_a.txt
timeout 300 python tools/refbench.py > gpurun_out/refbench_a.txt 2>&1; tail -30 gpurun_out/refbench_a.txt
{ echo "== default"; timeout 120 build/latbench 2000; echo "== FLB_SMALL=0"; FLB_SMALL=0 timeout 120 build/latbench 2000; } > gpurun_out/latbench_a.txt 2>&1; cat gpurun_out/latbench_a.txt
