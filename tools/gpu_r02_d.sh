#!/bin/bash
# round-2 call D (1 GPU): full -m gpu suite on the new kernels + their A/B switches, op benches, full bench.py, launch list
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest default exit $?"; tail -4 gpurun_out/pytest_gpu_d.log
FLB_ORIG_OCC=1 FLB_U8_DELTA_W8=slice FLB_U16_FILTER=warp FLB_SELECT=lane timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scan.py tests/test_golden.py -x -q -m gpu > gpurun_out/pytest_gpu_d2.log 2>&1; echo "pytest alt switches exit $?"; tail -4 gpurun_out/pytest_gpu_d2.log
echo "== filter u16 slice (default)"; timeout 300 python tools/opbench.py unpack_filter --types 16 2>&1 | tee gpurun_out/opbench_filter_u16_slice.txt
echo "== filter u16 warp"; FLB_U16_FILTER=warp timeout 300 python tools/opbench.py unpack_filter --types 16 2>&1 | tee gpurun_out/opbench_filter_u16_warp.txt
echo "== orig chains FLB_ORIG_OCC=0"; FLB_ORIG_OCC=0 timeout 400 python tools/opbench.py undelta_pack_untranspose,transpose_delta_pack --types 16,32,64 2>&1 | tee gpurun_out/opbench_orig_occ0.txt
echo "== orig chains FLB_ORIG_OCC=1"; FLB_ORIG_OCC=1 timeout 400 python tools/opbench.py undelta_pack_untranspose --types 32,64 2>&1 | tee gpurun_out/opbench_orig_occ1.txt
echo "== undelta_pack u8 warp W8"; timeout 300 python tools/opbench.py undelta_pack --types 8 2>&1 | tee gpurun_out/opbench_u8_delta_warp.txt
echo "== undelta_pack u8 slice W8"; FLB_U8_DELTA_W8=slice timeout 300 python tools/opbench.py undelta_pack --types 8 2>&1 | tee gpurun_out/opbench_u8_delta_slice.txt
echo "== select defaults"; timeout 300 python tools/opbench.py unpack_select_25pct 2>&1 | tee gpurun_out/opbench_select_defaults.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench exit $?"; tail -3 gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_d.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',d['value'],'frac',r['frac'],'min',r['min_frac_over_widths'],r['min_frac_width'])
print('e2e',{k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step','frac_of_ceiling','d2h_GBps_per_gpu')}, 'ceiling', d['e2e']['link_ceiling']['value'])
o=r['other']; print('min_frac_over_ops',o.get('min_frac_over_ops'),o['ops']['min_frac_op'] if 'ops' in o else None)
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu --no-other > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:unpack_warp_kernel -s 2 -c 1 -f -o /tmp/prof_unpack_u32_w16 python tools/ncu_one.py unpack 32 16 > gpurun_out/ncu_unpack_w16.log 2>&1; echo "ncu unpack exit $?"
ncu -i /tmp/prof_unpack_u32_w16.ncu-rep --page raw --csv > gpurun_out/ncu_raw_unpack_u32_w16.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:filter_u16_slice_kernel -s 2 -c 1 -f -o /tmp/prof_filter_u16_slice_w9 python tools/ncu_one.py unpack_filter 16 9 > gpurun_out/ncu_filter_u16_slice.log 2>&1; echo "ncu filter u16 slice exit $?"
ncu -i /tmp/prof_filter_u16_slice_w9.ncu-rep --page raw --csv > gpurun_out/ncu_raw_filter_u16_slice_w9.csv 2>/dev/null
timeout 120 build/test_traits > gpurun_out/test_traits.log 2>&1; echo "test_traits exit $?"; tail -2 gpurun_out/test_traits.log
du -sh gpurun_out
