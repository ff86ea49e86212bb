#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r01.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
for w in 1 16 32; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:unpack_warp_kernel -s 3 -c 1 -f \
      -o /tmp/prof_unpack_u32_w$w build/kbench/kb_u32 32 unpackT 20 1 $w $w > gpurun_out/ncu_w$w.log 2>&1
  ncu -i /tmp/prof_unpack_u32_w$w.ncu-rep --page raw --csv > gpurun_out/ncu_raw_unpack_u32_w$w.csv 2>/dev/null
  ncu -i /tmp/prof_unpack_u32_w$w.ncu-rep --page source --csv > gpurun_out/ncu_source_unpack_u32_w$w.csv 2>/dev/null
done
timeout 900 python tools/opbench.py > gpurun_out/opbench_r01.log 2>&1; tail -3 gpurun_out/opbench_r01.log
