#!/usr/bin/env python
"""tools/ncu_digest.py NAME [NAME...] — one compact table from gpurun_out/ncu_raw_<NAME>.csv captures
(`ncu --set full --clock-control none ... --page raw --csv`): duration, DRAM / L2 / issue utilisation, registers,
occupancy, the top warp-stall reasons.  Output: markdown on stdout (redirect into profiles/)."""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = [
    ("time us", "gpu__time_duration.sum"),
    ("dram rd GB", "dram__bytes_read.sum"),
    ("dram wr GB", "dram__bytes_write.sum"),
    ("dram %pk", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 %pk", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1 %pk", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("alu %", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("inst M", "sm__inst_executed.sum"),
    ("regs", "launch__registers_per_thread"),
    ("warps act %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("smem/blk", "launch__shared_mem_per_block_dynamic"),
    ("L2 wr sect M", "lts__t_sectors_srcunit_tex_op_write.sum"),
    ("smem wavefronts M", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("bank confl ld M", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("bank confl st M", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
]


def conv(val, unit):
    try:
        v = float(val.replace(",", ""))
    except ValueError:
        return val
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit)
    return v * mult if mult else v


def load(name):
    rows = list(csv.reader(open(os.path.join(SRC, f"ncu_raw_{name}.csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    names = sys.argv[1:]
    print("| capture | " + " | ".join(k for k, _ in KEYS) + " | top stalls (per-warp %, issue-stall reasons) |")
    print("|---|" + "---|" * (len(KEYS) + 1))
    for name in names:
        d = load(name)
        cells = []
        for label, key in KEYS:
            if key not in d:
                cells.append("-")
                continue
            v = conv(*d[key])
            if isinstance(v, float):
                if "GB" in label:
                    v = f"{v / 1e9:.3f}"
                elif label.endswith(" M"):
                    v = f"{v / 1e6:.1f}"
                else:
                    v = f"{v:.1f}"
            cells.append(str(v))
        stalls = []
        for h, (v, u) in d.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(v.replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print(f"| {name} | " + " | ".join(cells) + " | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:4]) + " |")
        print(f"|  | kernel: `{d['Kernel Name'][0][:110]}` |" + " |" * len(KEYS))


if __name__ == "__main__":
    main()
