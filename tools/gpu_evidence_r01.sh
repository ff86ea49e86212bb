#!/bin/bash
# round-1 evidence pass (final library): tests, both bench arms, launch list,
# ncu --set full of the headline kernel (W=16) and the filter kernel (W=8), full op table
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_s2.json 2> gpurun_out/bench_ref_s2.err; echo "ref exit $?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; echo "bench exit $?"; tail -3 gpurun_out/bench_s2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s2.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:unpack_warp_kernel -s 2 -c 1 -f -o /tmp/prof_unpack_u32_w16 \
    python tools/ncu_one.py unpack 32 16 > gpurun_out/ncu_unpack_w16.log 2>&1; echo "ncu unpack exit $?"
ncu -i /tmp/prof_unpack_u32_w16.ncu-rep --page raw --csv > gpurun_out/ncu_raw_unpack_u32_w16.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_warp_kernel -s 2 -c 1 -f -o /tmp/prof_filter_u32_w8 \
    python tools/ncu_one.py unpack_filter 32 8 > gpurun_out/ncu_filter_w8.log 2>&1; echo "ncu filter exit $?"
ncu -i /tmp/prof_filter_u32_w8.ncu-rep --page raw --csv > gpurun_out/ncu_raw_filter_u32_w8.csv 2>/dev/null
ncu -i /tmp/prof_filter_u32_w8.ncu-rep --page source --csv > gpurun_out/ncu_source_filter_u32_w8.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:select_warp_kernel -s 2 -c 1 -f -o /tmp/prof_select_u32_w8 \
    python tools/ncu_one.py unpack_select 32 8 > gpurun_out/ncu_select_w8.log 2>&1; echo "ncu select exit $?"
ncu -i /tmp/prof_select_u32_w8.ncu-rep --page raw --csv > gpurun_out/ncu_raw_select_u32_w8.csv 2>/dev/null
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -x > gpurun_out/sanitizer_memcheck_s2.txt 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/sanitizer_memcheck_s2.txt
timeout 1200 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py -q -m gpu -x -k "filter_every_width or select_every_width or pipeline or for_pack_auto or delta_filter" > gpurun_out/sanitizer_racecheck_s2.txt 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/sanitizer_racecheck_s2.txt
timeout 900 python tools/opbench.py > gpurun_out/opbench_s2.log 2>&1; echo "opbench exit $?"; tail -3 gpurun_out/opbench_s2.log
du -sh gpurun_out
