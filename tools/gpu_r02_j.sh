#!/bin/bash
# round-2 call J (1 GPU): select_warp2_kernel with early loads — parity, A/B (NB 1/2), ncu of u32 W=8 / u8 W=5 / u16 W=9
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
for mode in "FLB_SELECT=warp2" "FLB_SELECT=warp2 FLB_SELECT_NB=2"; do
  env $mode timeout 600 python -m pytest tests/test_gpu_scan.py -x -q -m gpu -k select > gpurun_out/pytest_gpu_j_sel.log 2>&1; echo "pytest [$mode] exit $?"; tail -2 gpurun_out/pytest_gpu_j_sel.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest parity exit $?"; tail -2 gpurun_out/pytest_gpu_j.log
echo "== warp2 nb1"; FLB_SELECT=warp2 timeout 300 python tools/opbench.py unpack_select_25pct 2>&1 | tee gpurun_out/opbench_select_j_warp2.txt
echo "== warp2 nb2"; FLB_SELECT=warp2 FLB_SELECT_NB=2 timeout 300 python tools/opbench.py unpack_select_25pct 2>&1 | tee gpurun_out/opbench_select_j_warp2_nb2.txt
cap() {  # name, kernel regex, op, T, W
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o /tmp/prof_$1 python tools/ncu_one.py $3 $4 $5 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
}
export FLB_SELECT=warp2
cap select_w2_u32_w8 select_warp2 unpack_select 32 8
ncu -i /tmp/prof_select_w2_u32_w8.ncu-rep --page source --csv > gpurun_out/ncu_source_select_w2_u32_w8.csv 2>/dev/null
cap select_w2_u8_w5 select_warp2 unpack_select 8 5
cap select_w2_u16_w9 select_warp2 unpack_select 16 9
cap select_w2_u64_w16 select_warp2 unpack_select 64 16
python tools/ncu_digest.py select_w2_u32_w8 select_w2_u8_w5 select_w2_u16_w9 select_w2_u64_w16 > gpurun_out/ncu_digest_j.md; cat gpurun_out/ncu_digest_j.md
