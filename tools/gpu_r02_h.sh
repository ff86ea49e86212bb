#!/bin/bash
# round-2 call H (1 GPU): re-verification after the container was re-created: full GPU suite, smoke, both bench arms
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_h.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench exit $?"; tail -2 gpurun_out/bench_h.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_h.json 2> gpurun_out/bench_ref_h.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_h.json').read().strip().splitlines()[-1])
e=d['e2e']; o=d['roofline']['other']
print('value',d['value'],'frac',d['roofline']['frac'],'clocks',d['clocks'])
print('e2e',e['value'],'ceiling',e['link_ceiling']['value'],e['frac_of_ceiling'])
print('min_frac_over_ops',o['ops']['min_frac_over_ops'],o['ops']['min_frac_op'])
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'])
r=json.loads(open('gpurun_out/bench_ref_h.json').read().strip().splitlines()[-1]); print('ref',r['value'],r['cpu_baseline'])
PY
