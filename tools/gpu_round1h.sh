#!/bin/bash
# warp-block pack/delta + tile transpose: parity + per-op table; PCIe/NUMA probe for the e2e path
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -12 gpurun_out/topo.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/opbench.py > gpurun_out/opbench_r1h.log 2>&1; cat gpurun_out/opbench_r1h.log
timeout 300 python - <<'PY' > gpurun_out/pcie_probe.log 2>&1
import os, time, torch, numpy as np, sys
sys.path.insert(0, os.getcwd())
import fastlanes_b200 as fl
def probe(tag):
    n = 1 << 30
    h = fl.pinned_empty(n, np.uint8); h[:] = 1
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    ht = torch.from_numpy(h)
    for name, fn in (("H2D", lambda: d.copy_(ht, non_blocking=True)), ("D2H", lambda: ht.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize()
        print(tag, name, round(5 * n / (time.perf_counter() - t) / 1e9, 1), "GB/s", flush=True)
    s2 = torch.cuda.Stream()
    t = time.perf_counter()
    for _ in range(5):
        d.copy_(ht, non_blocking=True)
        with torch.cuda.stream(s2): ht2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    print(tag, "bidir each", round(5 * n / (time.perf_counter() - t) / 1e9, 1), "GB/s", flush=True)
h2 = fl.pinned_empty(1 << 30, np.uint8); ht2 = torch.from_numpy(h2); d2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
print("affinity", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
probe("default")
for lo, hi in ((0, 32), (32, 64)):
    os.sched_setaffinity(0, set(range(lo, hi)) | set(range(lo + 64, hi + 64)))
    probe(f"cpus{lo}-{hi}")
PY
cat gpurun_out/pcie_probe.log
