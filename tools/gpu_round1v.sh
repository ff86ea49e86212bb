#!/bin/bash
# round 1, session 2, call 1: scan kernels parity + refactor regression + op table of the scan ops + quick bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu_v.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_scan.py -q -m gpu -x > gpurun_out/pytest_scan.log 2>&1; echo "scan pytest exit $?"; tail -15 gpurun_out/pytest_scan.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "parity pytest exit $?"; tail -3 gpurun_out/pytest_parity.log
timeout 600 python tools/opbench.py unpack_filter,unpack_select_25pct,unpack > gpurun_out/opbench_scan.log 2>&1; echo "opbench exit $?"; grep -E "filter|select" gpurun_out/opbench_scan.log | head -50
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v.json'))
print({k:d[k] for k in ('value','gbps','ms_per_step','e2e','clocks')})
r=d['roofline']; print(r['achieved'],r['peak'],r['frac'],r['min_frac_over_widths'],r['min_frac_width'],r['peak_source'])
PY
