#!/bin/bash
# A/B: pipelined unpack (NB blocks per warp, 2-stage TMA ring) vs the shipped one-block-per-warp kernel, u32
mkdir -p gpurun_out
FLB_UNPACK_PIPE=4 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "unpack_every_width or for_family or ragged or empty" > gpurun_out/pytest_pipe.log 2>&1; echo "pytest pipe exit $?"; tail -3 gpurun_out/pytest_pipe.log
for nb in 0 2 4 8; do
  FLB_UNPACK_PIPE=$nb timeout 300 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/bench_pipe$nb.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pipe$nb.json")); pw=d["roofline"]["per_width"]
print("pipe=$nb value", d["value"], "frac", d["roofline"]["frac"], " ".join(f"{w}:{pw[w]['GBps']:.0f}" for w in pw))
PY
done
