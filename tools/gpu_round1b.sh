#!/bin/bash
# second GPU call: smoke, bench (both arms), ncu launch list + full capture, store-pattern microbench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err; tail -c 600 gpurun_out/bench_ref_r01.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_r01.json; tail -5 gpurun_out/bench_r01.err
timeout 300 build/kbench/wbench > gpurun_out/wbench_r1b.log 2>&1; cat gpurun_out/wbench_r1b.log
# ncu: launch list of the bench command (serialised, cold cache: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
# ncu --set full on the dominant kernel at three widths
for w in 1 16 32; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:unpack_kernel -s 3 -c 1 -f \
      -o gpurun_out/prof_unpack_u32_w$w build/kbench/kb_base 32 unpack 20 1 $w $w > gpurun_out/ncu_w$w.log 2>&1
done
ls -la gpurun_out
