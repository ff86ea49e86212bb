#!/bin/bash
# round-2 call AA (1 GPU): per-instruction ncu digests (tools/ncu_mine.py) of the kernels furthest below their roofline
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
cap() {  # name, kernel regex, op, T, W, log2 blocks
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o /tmp/prof_$1 python tools/ncu_one.py $3 $4 $5 $6 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/ncu_source_$1.csv 2>/dev/null
  python tools/ncu_mine.py $1 $6
}
cap dfilter_u32_w8 delta_filter_warp undelta_pack_filter 32 8 20
cap filter_u64_w33 filter_warp unpack_filter 64 33 19
cap unfor_u8_w1 unpack unfor_pack 8 1 22
cap udo_u16_w1 unpack_warp undelta_pack_untranspose 16 1 21
cap undelta_u8_w5 unpack undelta_pack 8 5 22
cap auto_u64_w33 pack_warp for_pack_auto 64 33 19
python tools/ncu_digest.py dfilter_u32_w8 filter_u64_w33 unfor_u8_w1 udo_u16_w1 undelta_u8_w5 auto_u64_w33 > gpurun_out/ncu_digest_aa.md; cat gpurun_out/ncu_digest_aa.md
