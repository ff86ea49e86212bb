#!/bin/bash
# round-2 call M (1 GPU): the single-kernel select dispatch — full suite, sanitizer over the select tests, bench with the
# scan table, ncu of the final select kernel
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_m.log
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py -q -m gpu -x -k select > gpurun_out/sanitizer_memcheck_select.txt 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/sanitizer_memcheck_select.txt; tail -3 gpurun_out/sanitizer_memcheck_select.txt
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py -q -m gpu -x -k "select_every_width or pipeline" > gpurun_out/sanitizer_racecheck_select.txt 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/sanitizer_racecheck_select.txt; tail -3 gpurun_out/sanitizer_racecheck_select.txt
timeout 600 python bench.py > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; echo "bench exit $?"; tail -2 gpurun_out/bench_m.err
timeout 300 python tools/opbench.py unpack_select_25pct 2>&1 | tee gpurun_out/opbench_select_m.txt
cap() {  # name, kernel regex, op, T, W
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o /tmp/prof_$1 python tools/ncu_one.py $3 $4 $5 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
}
cap select_final_u32_w8 select_warp unpack_select 32 8
ncu -i /tmp/prof_select_final_u32_w8.ncu-rep --page source --csv > gpurun_out/ncu_source_select_final_u32_w8.csv 2>/dev/null
cap select_final_u8_w5 select_warp unpack_select 8 5
cap select_final_u16_w9 select_warp unpack_select 16 9
cap select_final_u64_w16 select_warp unpack_select 64 16
python tools/ncu_digest.py select_final_u32_w8 select_final_u8_w5 select_final_u16_w9 select_final_u64_w16 > gpurun_out/ncu_digest_m.md; cat gpurun_out/ncu_digest_m.md
