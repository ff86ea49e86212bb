#!/bin/bash
# final round-1 evidence pass: tests, both bench arms, ncu launch list + full captures, op table
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/ncu_*
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err; echo "ref exit $?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r01.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
for w in 1 16 32; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:unpack_warp_kernel -s 3 -c 1 -f \
      -o /tmp/prof_unpack_u32_w$w build/kbench/kb_u32 32 unpackT 20 1 $w $w > gpurun_out/ncu_w$w.log 2>&1
  ncu -i /tmp/prof_unpack_u32_w$w.ncu-rep --page raw --csv > gpurun_out/ncu_raw_unpack_u32_w$w.csv 2>/dev/null
  ncu -i /tmp/prof_unpack_u32_w$w.ncu-rep --page source --csv > gpurun_out/ncu_source_unpack_u32_w$w.csv 2>/dev/null
done
timeout 600 ncu --set full --clock-control none -k regex:pack_warp_kernel -s 3 -c 1 -f -o /tmp/prof_pack_u32_w16 build/kbench/kb_u32 32 packT 20 1 16 16 > gpurun_out/ncu_pack.log 2>&1
ncu -i /tmp/prof_pack_u32_w16.ncu-rep --page raw --csv > gpurun_out/ncu_raw_pack_u32_w16.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:unpack_warp_kernel -s 3 -c 1 -f -o /tmp/prof_undelta_u32_w8 build/kbench/kb_u32 32 undelta_packB 20 1 8 8 > gpurun_out/ncu_undelta.log 2>&1
ncu -i /tmp/prof_undelta_u32_w8.ncu-rep --page raw --csv > gpurun_out/ncu_raw_undelta_pack_u32_w8.csv 2>/dev/null
timeout 900 python tools/opbench.py > gpurun_out/opbench_r01.log 2>&1; tail -5 gpurun_out/opbench_r01.log
du -sh gpurun_out
