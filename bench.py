#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout (rank 0).

Workload (configs[1]): u32 `BitPacking::unpack`, width sweep W = 1..32, 2^20 blocks (2^30 values) per
width per GPU.  One "step" = the 32 kernel launches of the sweep over device-resident packed input
(uniform-random bits generated on the device: every bit pattern is a valid packing of uniform-random
W-bit values) into a 4 GiB output buffer.  Every launch streams >= 4.1 GiB, far beyond the 126 MB L2.

  value      whole-job billion ints/s over N GPUs, inputs resident in HBM (CUDA events, max over ranks)
  e2e        the same sweep through the host-buffer C-ABI call fl_host_unpack_u32 (page-locked host buffers placed on
             the GPU's NUMA node; H2D + kernels + D2H inside the timed region), >= 3 steps, median step;
             e2e.link_ceiling = the same byte counts through the same copy pipeline with NO kernel
             (fl_host_copy_probe), all ranks concurrently; e2e.frac_of_ceiling = ceiling time / e2e time
  roofline   algorithmic bytes / event-timed kernel duration vs the measured HBM peak (median of K launches per width);
             roofline.other = the other BASELINE configs, driver-run: configs[2] u64 pack + unpack, configs[3] fused
             undelta_pack vs unfused unpack + undelta (shape of benches/delta.rs:29-43), configs[4] the 2^26-block
             W=16 batch sharded over the ranks (strong scaling), and at N=1 the op x type x width table with
             min_frac_over_ops
  cpu_baseline  the CPU oracle (C++ restatement of the reference loops) on the box's host cores, same sweep, same size

`--impl reference` times the reference's CPU path instead (the crate cannot be built here: no Rust
toolchain, so it is the oracle port — see DESIGN.md) on all host threads, on the SAME config (2^20 blocks per width).

Multi-GPU (torchrun, one rank per GPU): blocks are independent, so every rank runs the same sweep
on its own shard with no data-path collective (weak scaling); NCCL carries only the barrier and the
max-over-ranks reduction of the timings.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_BITS = 32
WIDTHS = list(range(1, 33))
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def algorithmic_bytes_per_block(width: int) -> int:
    """unpack: 128*W bytes read + 128*T bytes written."""
    return 128 * (width + T_BITS)


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()  # exact PID we started
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


_CPU_BUFS = {}


def cpu_sweep(oracle, np, log2_blocks: int, threads: int, repeats: int):
    """The oracle's unpack over the same width sweep; returns (best seconds, ints).
    Buffers are created (and page-touched) once, outside the timed loop."""
    n = 1 << log2_blocks
    if log2_blocks not in _CPU_BUFS:
        rng = np.random.default_rng(42)
        # first-touch both buffers from the worker threads (NUMA-local pages), then fill the input
        packed = np.empty(n * 32 * 32, dtype=np.uint32)  # sized for W = 32
        out = np.empty(n * 1024, dtype=np.uint32)
        nt = oracle.hardware_threads()
        oracle.run_raw(32, oracle.OP_UNPACK, 0, n, None, packed, threads=nt)
        oracle.run_raw(32, oracle.OP_UNPACK, 0, n, None, out, threads=nt)
        step = 1 << 24
        for i in range(0, packed.size, step):
            packed[i:i + step] = rng.integers(0, 1 << 32, size=min(step, packed.size - i), dtype=np.uint32)
        _CPU_BUFS[log2_blocks] = (packed, out)
    packed, out = _CPU_BUFS[log2_blocks]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for w in WIDTHS:
            oracle.run_raw(32, oracle.OP_UNPACK, w, n, packed, out, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, n * 1024 * len(WIDTHS)


def cpu_filter_sweep(oracle, np, log2_blocks: int, threads: int, repeats: int):
    """CPU side of e2e.scan_filter: the oracle's unfor_pack + predicate loop over the same width sweep."""
    n = 1 << log2_blocks
    cpu_sweep(oracle, np, log2_blocks, threads, 1)  # creates the buffers
    packed, _ = _CPU_BUFS[log2_blocks]
    bitmap = np.empty(n * 128, dtype=np.uint8)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for w in WIDTHS:
            m = (1 << w) - 1
            oracle.unfor_filter(packed[: n * 32 * w], 0, w, m // 4, m // 2, n_blocks=n, threads=threads, out=bitmap)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, n * 1024 * len(WIDTHS)


def cgroup_cpu_limit():
    """CPUs this container may actually use: cgroup v2 cpu.max / v1 cfs quota (None = unlimited)."""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            return max(1, int(q) // int(per))
    except Exception:
        pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0:
            return max(1, q // per)
    except Exception:
        pass
    return None


def best_thread_count(oracle, np):
    """'All the host threads it can use': a container can expose more logical CPUs than its cgroup CPU quota
    (the GPU boxes of this pool: 128 visible, quota 16); threads beyond the quota only get throttled.  Probe
    thread counts up to min(logical CPUs, 2 x quota) on runs long enough to hit the throttle, keep the fastest."""
    hw = oracle.hardware_threads()
    quota = cgroup_cpu_limit()
    cap = hw if quota is None else min(hw, 2 * quota)
    cands = sorted({t for t in (1, 2, 4, 8, 16, 32, 64, 128, 256, quota or hw, cap) if 1 <= t <= cap})
    best_t, best_rate = 1, 0.0
    for t in cands:
        cpu_sweep(oracle, np, 16, t, 1)
        dt, ints = cpu_sweep(oracle, np, 16, t, 3)
        if ints / dt > best_rate * 1.03:
            best_t, best_rate = t, ints / dt
    return best_t, hw, quota


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port) on all host threads; rank 0 only.  Same config as the
    repo's arm: 2^20 blocks per width per step (4 GiB packed + 4 GiB output in host memory)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np

    from oracle import fl_oracle as oracle

    threads, hw, quota = best_thread_count(oracle, np)
    lg = args.cpu_log2_blocks
    for _ in range(max(1, min(args.warmup, 3))):
        cpu_sweep(oracle, np, lg, threads, 1)
    t0 = time.perf_counter()
    ints = 0
    for _ in range(args.steps):
        dt, n_ints = cpu_sweep(oracle, np, lg, threads, 1)
        ints += n_ints
    total = time.perf_counter() - t0
    gints = ints / total / 1e9
    sample = (f"u32 unpack W=1..32, 2^{lg} blocks per width per step (host memory), {oracle.isa()}, "
              f"{threads} threads (fastest of the probed counts; {hw} logical CPUs visible, cgroup CPU quota {quota})")
    line = {
        "impl": "reference", "metric": "u32 unpack width sweep throughput", "value": gints, "unit": "Gint/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "configs[1]: u32 unpack, width sweep W=1..32, 2^%d blocks per width per GPU" % lg,
                   "blocks_per_width_per_gpu": 1 << lg, "widths": "1..32"},
        "cpu_baseline": {"value": gints, "unit": "Gint/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gints, "unit": "Gint/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference crate is Rust nightly-2024-06-19 (no toolchain in this image): timed the C++ restatement of its loops",
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
class Dev:
    """What every measurement below needs: torch, the raw C-ABI loader, the launching stream, the rank layout."""

    def __init__(self, torch, _lib, dist, dev, world, rank, local_rank):
        self.torch, self._lib, self.dist, self.dev = torch, _lib, dist, dev
        self.world, self.rank, self.local_rank = world, rank, local_rank
        self.stream = torch.cuda.current_stream()
        self.sp = self.stream.cuda_stream

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        from fastlanes_b200.shard import max_over_ranks

        return max_over_ranks(v, self.dist, self.dev)

    def event_times(self, fn, k):
        """k launches of fn, each bracketed by events on the launching stream -> list of ms."""
        evs = []
        for _ in range(k):
            a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            a.record(self.stream); fn(); b.record(self.stream)
            evs.append((a, b))
        self.torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    def call(self, name, tbits, *args):
        st = self._lib.fn(name, tbits)(*args)
        if st != 0:
            self._lib.check(st)


def rec(ms_list, n_blocks, bytes_per_block, peak):
    ms = statistics.median(ms_list)
    gbps = n_blocks * bytes_per_block / (ms * 1e-3) / 1e9
    return {"us": round(ms * 1e3, 1), "GBps": round(gbps, 1), "Gints": round(n_blocks * 1024 / (ms * 1e-3) / 1e9, 1),
            "frac": round(gbps / peak, 4)}


def other_configs(D: Dev, packed, out, peak, steps, log2_total):
    """The BASELINE configs besides the headline sweep, each as its own small record (device-resident, CUDA events,
    median of K launches; every launch streams > 4 GiB)."""
    torch = D.torch
    K = max(5, min(steps, 20))
    res = {}
    P, U = packed.data_ptr(), out.data_ptr()
    # configs[2]: u64 pack + unpack, W in {1,17,33,48,64}; 2^19 blocks = 4 GiB unpacked (`out` reinterpreted as u64)
    n64 = out.numel() * 4 // 8192
    u64 = {}
    for w in (1, 17, 33, 48, 64):
        D.call("fl_unpack", 64, w, n64, P, U, D.sp)
        up = rec(D.event_times(lambda: D.call("fl_unpack", 64, w, n64, P, U, D.sp), K), n64, 128 * (w + 64), peak)
        pk = rec(D.event_times(lambda: D.call("fl_pack", 64, w, n64, U, P, D.sp), K), n64, 128 * (w + 64), peak)
        u64[str(w)] = {"unpack": up, "pack": pk}
    res["u64_pack_unpack"] = {"config": "configs[2]: u64 pack + unpack, W in {1,17,33,48,64}, 2^%d blocks (4 GiB unpacked)" % (n64.bit_length() - 1),
                              "per_width": u64, "min_frac": min(min(v["unpack"]["frac"], v["pack"]["frac"]) for v in u64.values())}
    # regenerate the packed bits the u64 pack overwrote (the headline sweep has already been timed; parity of the e2e
    # check later needs random input, not this output) — cheap, on device
    gen = torch.Generator(device=D.dev); gen.manual_seed(4242 + D.rank)
    for i in range(0, packed.numel(), 1 << 26):
        packed[i:i + (1 << 26)].random_(-(1 << 31), (1 << 31) - 1, generator=gen)
    # configs[3]: fused Delta + BitPack decode u32 W=8 (src/delta.rs:48-63) vs unfused unpack + undelta (benches/delta.rs:29-43)
    n = out.numel() // 1024
    base = torch.empty(n * 32, dtype=torch.int32, device=D.dev)
    base.random_(-(1 << 31), (1 << 31) - 1, generator=gen)
    tmp = torch.empty_like(out)
    B, TMP = base.data_ptr(), tmp.data_ptr()
    fused = rec(D.event_times(lambda: D.call("fl_undelta_pack", 32, 8, n, P, B, U, D.sp), K + 3)[3:], n, 128 * (8 + 32 + 1), peak)

    def unfused():
        D.call("fl_unpack", 32, 8, n, P, TMP, D.sp)
        D.call("fl_undelta", 32, n, TMP, B, U, D.sp)

    unf = rec(D.event_times(unfused, K + 3)[3:], n, 128 * (8 + 32) + 128 * (2 * 32 + 1), peak)
    # fused == unfused on the full buffer (device-side identity; the oracle comparison is in tests/test_gpu_configs.py)
    D.call("fl_undelta_pack", 32, 8, n, P, B, TMP, D.sp)
    torch.cuda.synchronize()
    same = bool(torch.equal(tmp, out))
    res["fused_delta_u32_w8"] = {"config": "configs[3]: fused undelta_pack u32 W=8, 2^%d blocks, vs unfused unpack + undelta" % (n.bit_length() - 1),
                                 "fused": fused, "unfused": unf, "speedup": round(unf["us"] / fused["us"], 3),
                                 "fused_equals_unfused": same}
    del tmp, base
    torch.cuda.empty_cache()
    # configs[4]: batched u32 W=16 unpack, 2^log2_total blocks TOTAL, contiguous shards over the ranks, waves of 2^22
    # blocks (8 GiB packed + 16 GiB out per wave >> L2) — strong scaling.  T(1) is measured in the same run: every rank
    # also runs the WHOLE batch alone (ranks do not interact), so efficiency = T1 / (N * TN) needs no second job.
    from fastlanes_b200.shard import block_shard, waves

    W, wave_blocks = 16, 1 << 22
    total = 1 << log2_total
    b0, b1 = block_shard(total, D.rank, D.world)
    wb = min(wave_blocks, total)
    p16 = torch.empty(wb * 32 * W, dtype=torch.int32, device=D.dev)
    for i in range(0, p16.numel(), 1 << 26):
        p16[i:i + (1 << 26)].random_(-(1 << 31), (1 << 31) - 1, generator=gen)
    o16 = torch.empty(wb * 1024, dtype=torch.int32, device=D.dev)

    def run_waves(n_mine):
        for _, nb in waves(n_mine, wb):
            D.call("fl_unpack", 32, W, nb, p16.data_ptr(), o16.data_ptr(), D.sp)

    def timed(n_mine, reps):
        run_waves(min(n_mine, wb))  # warm
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        e0.record(D.stream)
        for _ in range(reps):
            run_waves(n_mine)
        e1.record(D.stream)
        D.barrier()
        return D.max_over_ranks(e0.elapsed_time(e1)) / reps

    reps = 3
    t_n = timed(b1 - b0, reps)
    t_1 = timed(total, reps) if D.world > 1 else t_n
    bytes_total = total * algorithmic_bytes_per_block(W)
    res["sharded_batch_u32_w16"] = {
        "config": "configs[4]: batched u32 W=16 unpack, 2^%d blocks total, contiguous shards over %d GPU(s), waves of 2^22 blocks" % (log2_total, D.world),
        "scaling": "strong", "ms": round(t_n, 3), "Gints": round(total * 1024 / (t_n * 1e-3) / 1e9, 1),
        "GBps": round(bytes_total / (t_n * 1e-3) / 1e9, 1), "frac_per_gpu": round(bytes_total / (t_n * 1e-3) / 1e9 / D.world / peak, 4),
        "ms_one_gpu_whole_batch": round(t_1, 3), "strong_scaling_efficiency": round(t_1 / (D.world * t_n), 4),
        "launches_per_gpu": len(list(waves(b1 - b0, wb)))}
    del p16, o16
    torch.cuda.empty_cache()
    # Result check of the SHARDED path (not only its speed): a deterministic 2^18-block W=16 column defined by block
    # index alone; every rank decodes its contiguous shard, the wrapping int64 sums are all-reduced (the only collective
    # that ever touches decoded data: 8 bytes per rank) and must equal the sum of the whole column decoded by one GPU.
    # Sampled blocks at the shard boundaries go to the host for the oracle comparison in the cpu_baseline leg.
    nv = 1 << 18
    v0, v1 = block_shard(nv, D.rank, D.world)

    def column(bfirst, bend):  # packed words of blocks [bfirst, bend): an LCG of the global word index
        idx = torch.arange(bfirst * 32 * W, bend * 32 * W, dtype=torch.int64, device=D.dev)
        return ((idx * 6364136223846793005 + 1442695040888963407) >> 29).to(torch.int32)

    mine_p = column(v0, v1)
    mine_o = torch.empty((v1 - v0) * 1024, dtype=torch.int32, device=D.dev)
    if v1 > v0:
        D.call("fl_unpack", 32, W, v1 - v0, mine_p.data_ptr(), mine_o.data_ptr(), D.sp)
    from fastlanes_b200.shard import sum_over_ranks

    mask63 = (1 << 63) - 1
    part = int(mine_o.to(torch.int64).sum().item()) & mask63
    sharded_sum = sum_over_ranks(part, D.dist, D.dev) & mask63
    whole_p = column(0, nv)
    whole_o = torch.empty(nv * 1024, dtype=torch.int32, device=D.dev)
    D.call("fl_unpack", 32, W, nv, whole_p.data_ptr(), whole_o.data_ptr(), D.sp)
    whole_sum = int(whole_o.to(torch.int64).sum().item()) & mask63
    assert sharded_sum == whole_sum, "sharded decode: all-reduced checksum differs from the single-GPU decode"
    sample = sorted({0, nv - 1, *[block_shard(nv, r, D.world)[0] for r in range(D.world)], *[max(0, block_shard(nv, r, D.world)[1] - 1) for r in range(D.world)]})
    res["sharded_verify"] = {"config": "2^18-block u32 W=16 column, contiguous shards, all-reduced checksum of the decoded shards vs one GPU decoding the whole column",
                             "checksum_sharded": sharded_sum, "checksum_one_gpu": whole_sum, "match": sharded_sum == whole_sum,
                             "sampled_blocks": sample}
    res["_samples"] = {"width": W, "blocks": sample,
                       "packed": [whole_p[b * 32 * W:(b + 1) * 32 * W].cpu().numpy() for b in sample],
                       "decoded": [whole_o[b * 1024:(b + 1) * 1024].cpu().numpy() for b in sample]}
    del mine_p, mine_o, whole_p, whole_o
    torch.cuda.empty_cache()
    return res


def ops_table(D: Dev, peak):
    """Every bandwidth-type op x element type x 5 widths through the C ABI (what tools/opbench.py prints), as
    algorithmic GB/s; min_frac_over_ops is the minimum of achieved / measured peak over all of them."""
    torch = D.torch
    TDT = {8: torch.uint8, 16: torch.int16, 32: torch.int32, 64: torch.int64}
    table, worst = {}, (10.0, None)
    for tb in (8, 16, 32, 64):
        n = (1 << 31) // (128 * tb)  # 2 GiB unpacked per type
        unp = torch.empty(n * 1024, dtype=TDT[tb], device=D.dev); unp.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        pk = torch.empty(n * 1024, dtype=TDT[tb], device=D.dev); pk.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        base = torch.empty(n * (1024 // tb), dtype=TDT[tb], device=D.dev); base.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        aux = torch.empty(2 * n, dtype=TDT[tb], device=D.dev)
        U, P, B, A = unp.data_ptr(), pk.data_ptr(), base.data_ptr(), aux.data_ptr()
        ref = 12345 % (1 << tb)

        def put(op, w, bpb, fn):
            nonlocal worst
            fn()
            r = rec(D.event_times(fn, 5), n, bpb, peak)
            table.setdefault(op, {}).setdefault(f"u{tb}", {})[str(w)] = r["GBps"]
            if r["frac"] < worst[0]:
                worst = (r["frac"], f"{op} u{tb} W={w}")

        for w in sorted({1, tb // 4, tb // 2 + 1, tb - 3, tb}):
            put("unpack", w, 128 * (w + tb), lambda: D.call("fl_unpack", tb, w, n, P, U, D.sp))
            put("pack", w, 128 * (w + tb), lambda: D.call("fl_pack", tb, w, n, U, P, D.sp))
            put("unfor_pack", w, 128 * (w + tb), lambda: D.call("fl_unfor_pack", tb, w, n, P, ref, U, D.sp))
            put("for_pack", w, 128 * (w + tb), lambda: D.call("fl_for_pack", tb, w, n, U, ref, P, D.sp))
            put("for_pack_auto", w, 128 * (w + tb) + 2 * (tb // 8), lambda: D.call("fl_for_pack_auto", tb, w, n, U, A, A + n * (tb // 8), P, D.sp))
            put("undelta_pack", w, 128 * (w + tb + 1), lambda: D.call("fl_undelta_pack", tb, w, n, P, B, U, D.sp))
            put("undelta_pack_untranspose", w, 128 * (w + tb + 1), lambda: D.call("fl_undelta_pack_untranspose", tb, w, n, P, B, U, D.sp))
            put("transpose_delta_pack", w, 128 * (w + tb + 1), lambda: D.call("fl_transpose_delta_pack", tb, w, n, U, B, P, D.sp))
            put("unpack_cwida", w, 128 * (w + tb), lambda: D.call("fl_unpack_cwida", tb, w, n, P, U, D.sp))
            put("pack_cwida", w, 128 * (w + tb), lambda: D.call("fl_pack_cwida", tb, w, n, U, P, D.sp))
        put("delta", 0, 128 * (2 * tb + 1), lambda: D.call("fl_delta", tb, n, U, B, P, D.sp))
        put("undelta", 0, 128 * (2 * tb + 1), lambda: D.call("fl_undelta", tb, n, U, B, P, D.sp))
        put("transpose", 0, 256 * tb, lambda: D.call("fl_transpose", tb, n, U, P, D.sp))
        put("untranspose", 0, 256 * tb, lambda: D.call("fl_untranspose", tb, n, U, P, D.sp))
        put("block_minmax", 0, 128 * tb, lambda: D.call("fl_block_minmax", tb, n, U, A, A + n * (tb // 8), D.sp))
        del unp, pk, base, aux
        torch.cuda.empty_cache()
    return {"GBps": table, "min_frac_over_ops": worst[0], "min_frac_op": worst[1],
            "note": "algorithmic GB/s (SURVEY.md §8d byte formulas), 2 GiB unpacked per type, median of 5 launches, widths {1, T/4, T/2+1, T-3, T}"}


def scan_table(D: Dev, peak):
    """The fused scan entry points x element type x 5 widths: fused range filter, delta range filter
    and select at ~25 % selectivity, as algorithmic GB/s and G values/s.  Below W ~ T/2 these are instruction-bound by
    construction (1024 predicate evaluations or compactions per 128*W bytes), so they are reported next to the bandwidth
    ops, not folded into min_frac_over_ops."""
    torch = D.torch
    TDT = {8: torch.uint8, 16: torch.int16, 32: torch.int32, 64: torch.int64}
    table = {}
    for tb in (8, 16, 32, 64):
        n = (1 << 31) // (128 * tb)  # 2 GiB unpacked per type
        pk = torch.empty(n * 1024, dtype=TDT[tb], device=D.dev); pk.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        base = torch.empty(n * (1024 // tb), dtype=TDT[tb], device=D.dev); base.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        bm = torch.empty(n * 128, dtype=torch.uint8, device=D.dev)
        cnt = torch.empty(n, dtype=torch.int32, device=D.dev)
        P, B = pk.data_ptr(), base.data_ptr()
        full = (1 << tb) - 1

        def put(op, w, bpb, fn):
            fn()
            r = rec(D.event_times(fn, 5), n, bpb, peak)
            table.setdefault(op, {}).setdefault(f"u{tb}", {})[str(w)] = {"GBps": r["GBps"], "frac": r["frac"], "Gvalues": r["Gints"]}

        widths = sorted({1, tb // 4, tb // 2 + 1, tb - 3, tb})
        for w in widths:
            m = (1 << w) - 1
            put("unpack_filter", w, 128 * w + 128 + 4,
                lambda: D.call("fl_unpack_filter", tb, w, n, P, None, 0, m // 4, m // 2, bm.data_ptr(), cnt.data_ptr(), D.sp))
            put("undelta_pack_filter", w, 128 * w + 128 + 128 + 4,
                lambda: D.call("fl_undelta_pack_filter", tb, w, n, P, B, full // 4, full // 2, bm.data_ptr(), cnt.data_ptr(), D.sp))
        # select: a value-independent bitmap of density 1/4 and the exclusive prefix of its per-block counts
        bm.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        bm2 = bm.clone(); bm2.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        bm &= bm2
        del bm2
        c64 = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=D.dev)[bm.long()].view(n, 128).sum(1)
        offs = torch.cumsum(c64, 0) - c64
        total = int(c64.sum().item())
        sel = torch.empty(total + 16, dtype=TDT[tb], device=D.dev)
        for w in widths:
            put("unpack_select_25pct", w, 128 * w + 128 + 8 + (tb // 8) * total / n,
                lambda: D.call("fl_unpack_select", tb, w, n, P, None, 7, bm.data_ptr(), offs.data_ptr(), sel.data_ptr(), D.sp))
        del pk, base, bm, cnt, c64, offs, sel
        torch.cuda.empty_cache()
    return {"ops": table,
            "note": "algorithmic bytes per block: filter 128*W + 132, delta filter 128*W + 260, select 128*W + 136 + sizeof(T) * selected; "
                    "2 GiB unpacked per type, median of 5 launches; Gvalues = 1024 * blocks / time"}


def end_to_end(D: Dev, fl, np, packed, out, n_blocks, e2e_steps, launch):
    """The sweep through the host-buffer C ABI, its copy-only ceiling, and the same sweep as a host-buffer scan."""
    torch, _lib = D.torch, D._lib
    h_packed = fl.pinned_empty(n_blocks * 32 * 32, np.uint32)
    h_out = fl.pinned_empty(n_blocks * 1024, np.uint32)
    torch.cuda.synchronize()
    torch.from_numpy(h_packed.view(np.int32)).copy_(packed)  # device-generated bits -> host input, outside the timed region
    host_unpack = _lib.fn("fl_host_unpack", 32)
    probe = _lib.lib().fl_host_copy_probe
    ints_per_step = D.world * len(WIDTHS) * n_blocks * 1024

    def sweep(fn, k):
        """k separately timed steps (host wall clock around the synchronous calls), all ranks started together;
        returns (median over steps of the max-over-ranks step seconds, the per-step list)."""
        for w in (1, 16, 32):
            fn(w)  # warm-up: sizes the staging buffers, touches the pages
        steps = []
        for _ in range(k):
            D.barrier()
            t0 = time.perf_counter()
            for w in WIDTHS:
                fn(w)
            torch.cuda.synchronize()
            steps.append(D.max_over_ranks(time.perf_counter() - t0))
        return statistics.median(steps), steps

    def do_unpack(w):
        st = host_unpack(w, n_blocks, h_packed.ctypes.data, h_out.ctypes.data)
        if st != 0:
            _lib.check(st)

    def do_probe(w):
        st = probe(128 * w, 4096, n_blocks, h_packed.ctypes.data, h_out.ctypes.data)
        if st != 0:
            _lib.check(st)

    e2e_s, e2e_list = sweep(do_unpack, e2e_steps)
    # last call's result must equal the device result of the same width (before the probe overwrites h_out)
    check = torch.from_numpy(h_out.view(np.int32)[: 1 << 20]).to(D.dev)
    launch(32)
    torch.cuda.synchronize()
    assert torch.equal(check, out[: 1 << 20]), "e2e host path disagrees with the device path"
    ceil_s, ceil_list = sweep(do_probe, e2e_steps)
    h2d = sum(128 * w for w in WIDTHS) * n_blocks
    d2h = len(WIDTHS) * n_blocks * 4096
    e2e = {"value": round(ints_per_step / e2e_s / 1e9, 3), "unit": "Gint/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "ms_per_step": round(e2e_s * 1e3, 1), "ms_steps": [round(x * 1e3, 1) for x in e2e_list],
           "d2h_GBps_per_gpu": round(d2h / e2e_s / 1e9, 1),
           "api": "fl_host_unpack_u32 (page-locked host buffers from fl_host_alloc, chunked 3-stream H2D/kernel/D2H pipeline), per rank",
           "link_ceiling": {"value": round(ints_per_step / ceil_s / 1e9, 3), "unit": "Gint/s", "ms_per_step": round(ceil_s * 1e3, 1),
                            "ms_steps": [round(x * 1e3, 1) for x in ceil_list], "d2h_GBps_per_gpu": round(d2h / ceil_s / 1e9, 1),
                            "how": "fl_host_copy_probe: the same H2D and D2H byte counts per width through the same chunked pipeline with no kernel, all ranks concurrently"},
           "frac_of_ceiling": round(ceil_s / e2e_s, 4),
           "device_numa_node": _lib.lib().fl_device_numa_node(D.local_rank),
           "buffer_numa_node": [fl.buffer_node(h_out), fl.buffer_node(h_out, h_out.nbytes - 4096)],
           "timing": "host wall clock around the synchronous C-ABI calls (includes PCIe copies), barrier before every step, max over ranks, median step"}
    # ---- the same sweep as a SCAN through host buffers: fl_host_unpack_filter_u32 (fused decode + range predicate,
    # SURVEY.md §8f rank 2): only the 128-byte bitmap + count per block come back.  Extra key, one step.
    h_bitmap = h_counts = None
    alloc_failed = 0.0
    try:
        h_bitmap = fl.pinned_empty(n_blocks * 128, np.uint8)
        h_counts = fl.pinned_empty(n_blocks, np.uint32)
    except (fl.FastLanesError, MemoryError):
        alloc_failed = 1.0
    if D.max_over_ranks(alloc_failed) > 0:  # agreed through a collective: no rank waits in a barrier the others never reach
        e2e["scan_filter"] = {"value": None, "error": "pinned allocation for the bitmap failed on a rank"}
    else:
        host_filter = _lib.fn("fl_host_unpack_filter", 32)

        def do_filter(w, lo=None, hi=None):
            m = (1 << w) - 1
            st = host_filter(w, n_blocks, h_packed.ctypes.data, 0, m // 4 if lo is None else lo, m // 2 if hi is None else hi,
                             h_bitmap.ctypes.data, h_counts.ctypes.data)
            if st != 0:
                _lib.check(st)

        f_s, _ = sweep(do_filter, 1)
        launch(32)
        torch.cuda.synchronize()
        m32 = (1 << 32) - 1
        vals = out[: 1 << 20].cpu().numpy().view(np.uint32)
        want = np.packbits((vals >= np.uint32(m32 // 4)) & (vals <= np.uint32(m32 // 2)), bitorder="little")
        assert np.array_equal(h_bitmap[: want.size], want), "e2e host filter disagrees with the device unpack"
        e2e["scan_filter"] = {
            "value": round(ints_per_step / f_s / 1e9, 2), "unit": "Gint/s scanned", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": len(WIDTHS) * n_blocks * 132, "ms_per_step": round(f_s * 1e3, 1), "steps": 1,
            "api": "fl_host_unpack_filter_u32: range predicate lo<=v<=hi per width, bitmap + counts to page-locked host memory"}
    # ---- the reference's own throughput bench shape (benches/bitpacking.rs:67-98: u16, W = 3, 1024 blocks = 2 MiB) through
    # the host family on page-locked buffers: one launch on the caller's memory (the direct path of host_op).  Rank 0 only.
    if D.rank == 0:
        try:
            nb = 1024
            v16 = fl.pinned_empty(nb * 1024, np.uint16); v16[:] = (np.arange(nb * 1024) % 8).astype(np.uint16)
            p16 = fl.pinned_empty(nb * 192, np.uint16)
            u16 = fl.pinned_empty(nb * 1024, np.uint16)
            pack16, unpack16 = _lib.fn("fl_host_pack", 16), _lib.fn("fl_host_unpack", 16)

            def med_us(fn, k=60):
                for _ in range(5):
                    fn()
                ts = []
                for _ in range(k):
                    t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
                return statistics.median(ts) * 1e6

            tp = med_us(lambda: pack16(3, nb, v16.ctypes.data, p16.ctypes.data))
            tu = med_us(lambda: unpack16(3, nb, p16.ctypes.data, u16.ctypes.data))
            assert np.array_equal(u16, v16), "1024-block round trip through the host family"
            e2e["ref_bench_shape"] = {
                "workload": "benches/bitpacking.rs:67-98: u16 W=3, 1024 blocks (2 MiB unpacked), page-locked host buffers",
                "compress_us": round(tp, 1), "decompress_us": round(tu, 1),
                "decompress_GBps_unpacked": round(nb * 2048 / tu / 1e3, 2),
                "api": "fl_host_pack_u16 / fl_host_unpack_u16: one launch on the caller's page-locked memory (direct path), host wall clock, median of 60 calls"}
            del v16, p16, u16
        except (fl.FastLanesError, MemoryError) as exc:  # reported, never fatal for the headline
            e2e["ref_bench_shape"] = {"error": str(exc)}
    del h_packed, h_out, h_bitmap, h_counts
    return e2e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-blocks", type=int, default=20, help="blocks per width per GPU (configs[1]: 20)")
    ap.add_argument("--e2e-steps", type=int, default=3, help="timed end-to-end (host buffer) steps, median reported; 0 disables")
    ap.add_argument("--cpu-log2-blocks", type=int, default=20, help="CPU arm: blocks per width (20 = the repo arm's config)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip roofline.other (configs[2..4])")
    ap.add_argument("--no-ops", action="store_true", help="skip the op x type table (it only runs at N=1)")
    ap.add_argument("--log2-total-blocks", type=int, default=26, help="configs[4]: total blocks of the sharded batch")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    import fastlanes_b200 as fl
    from fastlanes_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    D = Dev(torch, _lib, dist, dev, world, rank, local_rank)

    n_blocks = 1 << args.log2_blocks
    # packed input sized for W = 32 (each width reads its own 128*W*n_blocks-byte prefix); 4 GiB output
    packed = torch.empty(n_blocks * 32 * 32, dtype=torch.int32, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(42 + rank)
    chunk = 1 << 26
    for i in range(0, packed.numel(), chunk):
        packed[i:i + chunk].random_(-(1 << 31), (1 << 31) - 1, generator=gen)
    out = torch.empty(n_blocks * 1024, dtype=torch.int32, device=dev)
    unpack = _lib.fn("fl_unpack", 32)
    stream, sp = D.stream, D.sp

    def launch(w):
        st = unpack(w, n_blocks, packed.data_ptr(), out.data_ptr(), sp)
        if st != 0:
            _lib.check(st)

    def step():
        for w in WIDTHS:
            launch(w)

    warm = max(3, args.warmup)
    for _ in range(warm):
        step()
    D.barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # ---- timed region: EXACTLY K steps, events on the launching stream ------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    t_mark0 = sampler.mark()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    D.barrier()
    t_mark1 = sampler.mark()
    total_ms = e0.elapsed_time(e1)
    launches = args.steps * len(WIDTHS)

    # ---- per-width kernel durations (roofline): ONE pass, K launches per width, events around each launch, median ----
    per_w = {w: D.event_times(lambda: launch(w), args.steps) for w in WIDTHS}
    if rank == 0:
        time.sleep(0.2)
        sampler.stop()

    total_ms = D.max_over_ranks(total_ms)
    ints_per_step = world * len(WIDTHS) * n_blocks * 1024
    value = ints_per_step * args.steps / (total_ms * 1e-3) / 1e9
    bytes_per_step_rank = sum(algorithmic_bytes_per_block(w) for w in WIDTHS) * n_blocks
    gbps = world * bytes_per_step_rank * args.steps / (total_ms * 1e-3) / 1e9

    peak, peak_src = load_peak()
    med = {w: statistics.median(per_w[w]) for w in WIDTHS}
    kern_ms = sum(med.values())
    achieved = bytes_per_step_rank / (kern_ms * 1e-3) / 1e9
    per_width = {str(w): {"us": round(med[w] * 1e3, 1), "us_mean": round(statistics.mean(per_w[w]) * 1e3, 1),
                          "GBps": round(algorithmic_bytes_per_block(w) * n_blocks / (med[w] * 1e-3) / 1e9, 1),
                          "Gints": round(n_blocks * 1024 / (med[w] * 1e-3) / 1e9, 1)} for w in WIDTHS}
    worst = min(WIDTHS, key=lambda w: per_width[str(w)]["GBps"])
    roofline = {
        "bound": "hbm", "kernel": "flb::unpack_warp_kernel<uint32_t, W, UOP_PLAIN, TMA> (W=1..32, one launch per width)",
        "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
        "peak_source": peak_src, "traffic": None,
        "algorithmic_bytes_per_launch": "128*(W+32) bytes/block * 2^%d blocks" % args.log2_blocks,
        "timing": "CUDA events around each launch on the launching stream; per width the MEDIAN of K launches of one pass",
        "frac_whole_timed_region": round(bytes_per_step_rank * args.steps / (total_ms * 1e-3) / 1e9 / peak, 4),
        "min_frac_over_widths": round(per_width[str(worst)]["GBps"] / peak, 4), "min_frac_width": worst,
        "per_width": per_width,
    }
    for name in ("traffic_r02.json", "traffic_r01.json"):
        tr = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tr):
            try:
                roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch_w16")
                roofline["traffic_source"] = "profiles/" + name + " (ncu --set full, W=16 launch: dram__bytes_read.sum + dram__bytes_write.sum)"
                break
            except Exception:
                pass

    other, samples = None, None
    if not args.no_other:
        other = other_configs(D, packed, out, peak, args.steps, args.log2_total_blocks)
        samples = other.pop("_samples")
        if world == 1 and not args.no_ops:
            other["ops"] = ops_table(D, peak)
            other["min_frac_over_ops"] = other["ops"]["min_frac_over_ops"]
            other["scan_ops"] = scan_table(D, peak)
    roofline["other"] = other

    # ---- end to end through the host-buffer C-ABI (page-locked host memory) ------------------------
    e2e = None
    if args.e2e_steps > 0:
        e2e = end_to_end(D, fl, np, packed, out, n_blocks, max(3, args.e2e_steps), launch)

    cpu = None
    if rank == 0 and not args.no_cpu:
        from oracle import fl_oracle as oracle

        threads, hw, quota = best_thread_count(oracle, np)
        lg = args.cpu_log2_blocks
        cpu_sweep(oracle, np, lg, threads, 1)
        dt, ints = cpu_sweep(oracle, np, lg, threads, 3)
        dt1, ints1 = cpu_sweep(oracle, np, 13, 1, 2)
        fdt, fints = cpu_filter_sweep(oracle, np, lg, threads, 2)
        # the reference's throughput bench shape on one thread, as criterion runs it (benches/bitpacking.rs:67-98)
        rb_v = (np.arange(1024 * 1024) % 8).astype(np.uint16)
        rb_p = oracle.pack(rb_v, 3)
        rb_u = np.zeros_like(rb_v)
        rb_t = []
        for _ in range(25):
            t0 = time.perf_counter(); oracle.run_raw(16, oracle.OP_UNPACK, 3, 1024, rb_p, rb_u); rb_t.append(time.perf_counter() - t0)
        rb_us = statistics.median(rb_t[5:]) * 1e6
        if samples is not None:  # the oracle as checker of the sharded decode (sampled blocks at the shard boundaries)
            ok = all(np.array_equal(oracle.unpack(p.view(np.uint32), samples["width"], n_blocks=1), d.view(np.uint32))
                     for p, d in zip(samples["packed"], samples["decoded"]))
            other["sharded_verify"]["oracle_sampled_blocks_match"] = bool(ok)
            assert ok, "sharded decode disagrees with the oracle on sampled blocks"
        cpu = {"value": round(ints / dt / 1e9, 3), "unit": "Gint/s", "cores": threads, "kind": "port",
               "sample": f"u32 unpack W=1..32, 2^{lg} blocks per width (the whole config; best of 3 passes), host memory, {oracle.isa()}, {threads} threads (fastest probed; {hw} logical CPUs visible, cgroup CPU quota {quota})",
               "single_thread_Gints": round(ints1 / dt1 / 1e9, 3),
               "scan_filter_Gints": round(fints / fdt / 1e9, 3),
               "ref_bench_shape_decompress_us_1thread": round(rb_us, 1),
               "scan_filter_note": "same threads: unfor_pack into a 4 KiB per-thread scratch + range-predicate loop -> bitmap (what a user of the reference writes, README.md:40-41); compare with e2e.scan_filter",
               "note": "C++ restatement of the reference loops (the Rust crate cannot be built here); a reported baseline, not the target"}

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    clocks = sampler.summary(t_mark0, t_mark1)
    line = {
        "metric": "u32 unpack width sweep throughput", "value": round(value, 2), "unit": "Gint/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": round(total_ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "gbps": round(gbps, 1),
        "config": {"workload": "configs[1]: u32 unpack, width sweep W=1..32, 2^%d blocks per width per GPU" % args.log2_blocks,
                   "blocks_per_width_per_gpu": n_blocks, "widths": "1..32", "parallelism": f"block-sharded x{world}, no data-path collective",
                   "l2": "every launch streams >= 4.1 GiB (input prefix + 4 GiB output) >> 126 MB L2; no flush needed",
                   "input": "device-generated uniform-random bits, seed 42+rank"},
        "gpu_launches": launches, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
