#!/bin/bash
# round-2 call Z (1 GPU): final library — full suite, smoke, both bench arms; u64 pack family A/B (whole 16-byte shared loads)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_z.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_z.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== new"; timeout 300 python tools/opbench.py pack,for_pack,for_pack_auto --types 64 2>&1 | tee gpurun_out/opbench_pack_u64_new.txt
cp fastlanes_b200/lib/libfastlanes_b200.so /tmp/new.so; cp build/lib_old/libfastlanes_b200.so fastlanes_b200/lib/libfastlanes_b200.so
echo "== old"; timeout 300 python tools/opbench.py pack,for_pack,for_pack_auto --types 64 2>&1 | tee gpurun_out/opbench_pack_u64_old.txt
cp /tmp/new.so fastlanes_b200/lib/libfastlanes_b200.so
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_z.json 2> gpurun_out/bench_ref_z.err; echo "ref exit $?"
timeout 600 python bench.py > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err; echo "bench exit $?"; tail -2 gpurun_out/bench_z.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_z.json').read().strip().splitlines()[-1])
e=d['e2e']; o=d['roofline']['other']
print('value',d['value'],'frac',d['roofline']['frac'],'clocks',d['clocks'])
print('e2e',e['value'],'ceiling',e['link_ceiling']['value'],e['frac_of_ceiling'], e.get('ref_bench_shape',{}).get('decompress_us'))
print('min_frac_over_ops',o['ops']['min_frac_over_ops'],o['ops']['min_frac_op'])
t=o['ops']['GBps']; peak=d['roofline']['peak']
rows=sorted((g/peak,op,ty,w) for op,a in t.items() for ty,b in a.items() for w,g in b.items())
print('below .95:',[(round(r[0],3),)+r[1:] for r in rows if r[0]<0.95])
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'])
r=json.loads(open('gpurun_out/bench_ref_z.json').read().strip().splitlines()[-1]); print('ref',r['value'])
PY
