#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_scan.py tests/test_gpu_threads.py tests/test_gpu_cpp_traits.py -q -m gpu > gpurun_out/pytest_scan.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_scan.log
timeout 600 python tools/opbench.py unpack_filter --types 8,16 > gpurun_out/opbench_q.log 2>&1; echo "opbench exit $?"; cat gpurun_out/opbench_q.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; echo "bench exit $?"; tail -2 gpurun_out/bench_s2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s2.json')); r=d['roofline']
print(d['value'], d['gbps'], r['achieved'], r['frac'], r['min_frac_over_widths'], r['min_frac_width'], d['e2e']['value'], d['clocks'])
PY
