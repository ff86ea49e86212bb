#!/bin/bash
# round-2 call E (1 GPU): u16 original-order slice kernels (tests + A/B), u16 OCC variant, u8 filter gather, select after hoist
mkdir -p gpurun_out
FLB_U16_ORIG=slice timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_configs.py tests/test_gpu_host_path.py -x -q -m gpu > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest u16 slice exit $?"; tail -4 gpurun_out/pytest_gpu_e.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scan.py tests/test_golden.py -x -q -m gpu > gpurun_out/pytest_gpu_e2.log 2>&1; echo "pytest defaults exit $?"; tail -4 gpurun_out/pytest_gpu_e2.log
echo "== u16 orig chains: warp (default)"; timeout 300 python tools/opbench.py undelta_pack_untranspose,transpose_delta_pack --types 16 2>&1 | tee gpurun_out/opbench_u16_orig_warp.txt
echo "== u16 orig chains: warp + FLB_ORIG_OCC=2"; FLB_ORIG_OCC=2 timeout 300 python tools/opbench.py undelta_pack_untranspose --types 16 2>&1 | tee gpurun_out/opbench_u16_orig_occ2.txt
echo "== u16 orig chains: slice"; FLB_U16_ORIG=slice timeout 300 python tools/opbench.py undelta_pack_untranspose,transpose_delta_pack --types 16 2>&1 | tee gpurun_out/opbench_u16_orig_slice.txt
python - <<'PY'
# finer width sweep for the u16 slice-vs-warp decision (W = 1..8)
import os, subprocess, sys
code = r'''
import sys, statistics, torch
sys.path.insert(0, ".")
from fastlanes_b200 import _lib
tb = 16; n = (1 << 32) // (128 * tb)
U = torch.empty(n * 1024, dtype=torch.int16, device="cuda"); U.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
P = torch.empty(n * 1024, dtype=torch.int16, device="cuda"); P.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
B = torch.empty(n * 64, dtype=torch.int16, device="cuda"); B.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
sp = torch.cuda.current_stream().cuda_stream
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
for w in range(1, 17):
    d = t(lambda: _lib.fn("fl_undelta_pack_untranspose", tb)(w, n, P.data_ptr(), B.data_ptr(), U.data_ptr(), sp))
    e = t(lambda: _lib.fn("fl_transpose_delta_pack", tb)(w, n, U.data_ptr(), B.data_ptr(), P.data_ptr(), sp))
    gb = n * 128 * (w + tb + 1) / 1e6
    print(f"W={w:2d} undelta_pack_untranspose {d*1e3:8.1f} us {gb/d:7.1f} GB/s | transpose_delta_pack {e*1e3:8.1f} us {gb/e:7.1f} GB/s", flush=True)
'''
for mode in ("warp", "slice"):
    print("== u16 width sweep FLB_U16_ORIG=" + mode, flush=True)
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FLB_U16_ORIG=mode))
PY
echo "== filter u8 (new pass-bit gather)"; timeout 300 python tools/opbench.py unpack_filter --types 8 2>&1 | tee gpurun_out/opbench_filter_u8_new.txt
echo "== select u16/u32 after hoist"; timeout 300 python tools/opbench.py unpack_select_25pct --types 16,32 2>&1 | tee gpurun_out/opbench_select_hoist.txt
timeout 200 build/numa_probe 29 > gpurun_out/numa_probe_e.txt 2>&1; grep -E "dev0|pci" gpurun_out/numa_probe_e.txt
