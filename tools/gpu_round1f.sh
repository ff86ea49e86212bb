#!/bin/bash
# full gpu test tier, cgroup probe, bench (both arms), other-type kbench
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/ncu_*
{ echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; echo "nproc: $(nproc)"; echo "cpuset: $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null)"; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node\(s\)"; } > gpurun_out/host_info.txt 2>&1; cat gpurun_out/host_info.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err; echo "ref exit $?"; tail -c 500 gpurun_out/bench_ref_r01.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r01.err
timeout 600 python bench.py --workload scaling --steps 3 --warmup 3 > gpurun_out/bench_scaling_n1.json 2> gpurun_out/bench_scaling_n1.err; echo "scaling exit $?"; cat gpurun_out/bench_scaling_n1.json | cut -c1-400
K=build/kbench
{
timeout 300 $K/kb_base 64 unpackB 19 5
timeout 200 $K/kb_base 16 unpackB 21 5
timeout 200 $K/kb_base 8 unpackB 22 5
for t in 8 16 64; do timeout 100 $K/kb_base $t undelta_packB 20 5 5 9; done
timeout 300 $K/kb_base 64 pack 19 5 1 64
for t in 8 16 32 64; do timeout 60 $K/kb_base $t delta 20 5; timeout 60 $K/kb_base $t undelta 20 5; done
} > gpurun_out/kbench_r1f.log 2>&1
grep -v "^#" gpurun_out/kbench_r1f.log | awk 'NR%4==1' | head -60
