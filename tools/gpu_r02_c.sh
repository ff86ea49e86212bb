#!/bin/bash
# round-2 call C (1 GPU): select A/B (thread vs lane-per-word), for_pack_auto A/B, latency floor, ncu of the sub-0.9 kernels
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest (lane select, slice auto) exit $?"; tail -4 gpurun_out/pytest_gpu_c.log
FLB_SELECT=thread FLB_AUTO_SLICE=2 timeout 600 python -m pytest tests/test_gpu_scan.py -x -q -m gpu > gpurun_out/pytest_gpu_c2.log 2>&1; echo "pytest (thread select, slice auto all) exit $?"; tail -4 gpurun_out/pytest_gpu_c2.log
timeout 300 python tools/opbench.py unpack_select_25pct > gpurun_out/opbench_select_lane.txt 2>&1; echo "== select lane"; cat gpurun_out/opbench_select_lane.txt
for m in 0 1 2; do echo "== for_pack_auto FLB_AUTO_SLICE=$m"; FLB_AUTO_SLICE=$m timeout 300 python tools/opbench.py for_pack_auto --types 8,16 2>&1 | tee gpurun_out/opbench_auto_slice$m.txt; done
timeout 120 build/launch_floor > gpurun_out/launch_floor.txt 2>&1; cat gpurun_out/launch_floor.txt
timeout 120 build/latbench 2000 > gpurun_out/latbench_c.txt 2>&1; tail -8 gpurun_out/latbench_c.txt
cap() {  # name kernel-regex op tbits width
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o /tmp/prof_$1 python tools/ncu_one.py $3 $4 $5 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
}
cap select_lane_u32_w8 select_lane_kernel unpack_select 32 8
ncu -i /tmp/prof_select_lane_u32_w8.ncu-rep --page source --csv > gpurun_out/ncu_source_select_lane_u32_w8.csv 2>/dev/null
cap select_lane_u8_w5 select_lane_kernel unpack_select 8 5
cap auto_u8_w1 for_pack_auto for_pack_auto 8 1
cap undelta_orig_u32_w1 unpack_warp_kernel undelta_pack_untranspose 32 1
cap undelta_orig_u16_w1 unpack_warp_kernel undelta_pack_untranspose 16 1
cap undelta_u8_w8 unpack_warp_kernel undelta_pack 8 8
cap undelta_u16_w1 unpack_warp_kernel undelta_pack 16 1
cap filter_u16_w9 filter_warp_kernel unpack_filter 16 9
cap tdp_u64_w1 pack_warp_kernel transpose_delta_pack 64 1
FLB_SELECT=thread timeout 300 ncu --set full --clock-control none -k regex:select_warp_kernel -s 2 -c 1 -f -o /tmp/prof_select_thread_u32_w8 python tools/ncu_one.py unpack_select 32 8 > gpurun_out/ncu_select_thread.log 2>&1
ncu -i /tmp/prof_select_thread_u32_w8.ncu-rep --page raw --csv > gpurun_out/ncu_raw_select_thread_u32_w8.csv 2>/dev/null
du -sh gpurun_out
