#!/bin/bash
# third GPU call: bench JSON (small outputs), pattern microbench 2, ncu summaries exported to CSV on the box
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench exit $?"
timeout 300 build/kbench/wbench2 > gpurun_out/wbench2_r1c.log 2>&1; cat gpurun_out/wbench2_r1c.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
for w in 1 16 32; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:unpack_kernel -s 3 -c 1 -f \
      -o /tmp/prof_unpack_u32_w$w build/kbench/kb_u32 32 unpack 20 1 $w $w > gpurun_out/ncu_w$w.log 2>&1
  ncu -i /tmp/prof_unpack_u32_w$w.ncu-rep --page raw --csv > gpurun_out/ncu_raw_unpack_u32_w$w.csv 2>/dev/null
  ncu -i /tmp/prof_unpack_u32_w$w.ncu-rep --page details --csv > gpurun_out/ncu_details_unpack_u32_w$w.csv 2>/dev/null
  ncu -i /tmp/prof_unpack_u32_w$w.ncu-rep --page source --csv > gpurun_out/ncu_source_unpack_u32_w$w.csv 2>/dev/null
  ls -la /tmp/prof_unpack_u32_w$w.ncu-rep
done
cp /tmp/prof_unpack_u32_w16.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out
