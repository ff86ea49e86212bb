#!/bin/bash
# round-2 call U (1 GPU): final library after the clean rebuild — full suite, bench with e2e.ref_bench_shape
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_u.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_u.log
timeout 600 python bench.py > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err; echo "bench exit $?"; tail -2 gpurun_out/bench_u.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_u.json').read().strip().splitlines()[-1])
e=d['e2e']; o=d['roofline']['other']
print('value',d['value'],'frac',d['roofline']['frac'])
print('e2e',e['value'],'ceiling',e['link_ceiling']['value'],e['frac_of_ceiling'])
print('ref_bench_shape',e.get('ref_bench_shape'))
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'],d['cpu_baseline'].get('ref_bench_shape_decompress_us_1thread'))
print('min_frac_over_ops',o['ops']['min_frac_over_ops'],o['ops']['min_frac_op'])
PY
