#!/bin/bash
# round-2 call P (1 GPU): full suite + smoke + both bench arms with the final library of the session
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_p.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_p.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_p.json 2> gpurun_out/bench_ref_p.err; echo "ref exit $?"
timeout 600 python bench.py > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench exit $?"; tail -2 gpurun_out/bench_p.err
timeout 300 python tools/opbench.py unpack_filter,undelta_pack_filter,unpack_select_25pct 2>&1 | tee gpurun_out/opbench_scan_p.txt | tail -5
timeout 300 python tools/refbench.py 2>&1 | tee gpurun_out/refbench_p.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_p.json').read().strip().splitlines()[-1])
e=d['e2e']; o=d['roofline']['other']
print('value',d['value'],'frac',d['roofline']['frac'],'clocks',d['clocks'])
print('e2e',e['value'],'ceiling',e['link_ceiling']['value'],e['frac_of_ceiling'])
print('min_frac_over_ops',o['ops']['min_frac_over_ops'],o['ops']['min_frac_op'])
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'])
r=json.loads(open('gpurun_out/bench_ref_p.json').read().strip().splitlines()[-1]); print('ref',r['value'])
PY
