#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2>/dev/null; echo "ref exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json')); r=d['roofline']
print(d['value'], d['gbps'], r['achieved'], r['frac'], r['min_frac_over_widths'], r['min_frac_width'], d['e2e']['value'], d['e2e']['scan_filter']['value'], d['cpu_baseline']['value'], d['clocks'])
print(json.load(open('gpurun_out/bench_ref_final.json'))['value'])
PY
timeout 600 python tools/opbench.py > gpurun_out/opbench_final.log 2>&1; echo "opbench exit $?"
