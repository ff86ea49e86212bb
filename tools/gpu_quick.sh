#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_warp_kernel -s 2 -c 1 -f -o /tmp/prof_select_u32_w8 \
    python tools/ncu_one.py unpack_select 32 8 > gpurun_out/ncu_select_w8.log 2>&1; echo "ncu select exit $?"
ncu -i /tmp/prof_select_u32_w8.ncu-rep --page raw --csv > gpurun_out/ncu_raw_select_u32_w8.csv 2>/dev/null
ncu -i /tmp/prof_select_u32_w8.ncu-rep --page details --csv > gpurun_out/ncu_details_select_u32_w8.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:filter_warp_kernel -s 2 -c 1 -f -o /tmp/prof_filter_u32_w8 \
    python tools/ncu_one.py unpack_filter 32 8 > gpurun_out/ncu_filter_w8.log 2>&1; echo "ncu filter exit $?"
ncu -i /tmp/prof_filter_u32_w8.ncu-rep --page raw --csv > gpurun_out/ncu_raw_filter_u32_w8.csv 2>/dev/null
