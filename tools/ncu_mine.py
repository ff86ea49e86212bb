#!/usr/bin/env python
"""tools/ncu_mine.py NAME [LOG2_BLOCKS] — per-instruction digest of gpurun_out/ncu_source_<NAME>.csv (`ncu --page source --csv`):
shared-memory wavefronts vs ideal per opcode, global sectors vs ideal, issued instructions per block, and the instructions
holding the most warp-stall samples.  Finds what the summary metrics hide (e.g. 16-byte shared loads narrowed by ptxas into
bank-conflicting 4-byte loads, profiles/ncu_r02_u64_orig.md)."""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    name = sys.argv[1]
    nblk = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
    rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"ncu_source_{name}.csv"))))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]

    def num(r, key):
        try:
            return int(r[ci[key]])
        except (ValueError, KeyError):
            return 0

    print(f"## {name}: {rows[0][1][:100]}")
    print(f"issued instructions per block: {sum(num(r, 'Instructions Executed') for r in data) / nblk:.1f}")
    agg = {}
    for r in data:
        w = num(r, "L1 Wavefronts Shared")
        if w:
            toks = r[ci["Source"]].split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            a = agg.setdefault(op, [0, 0, 0])
            a[0] += w; a[1] += num(r, "L1 Wavefronts Shared Ideal"); a[2] += num(r, "Instructions Executed")
    for op, a in sorted(agg.items()):
        print(f"  shared {op:9s} executed/block {a[2] / nblk:6.1f}  wavefronts/block {a[0] / nblk:6.1f}  ideal {a[1] / nblk:6.1f}")
    gs, gi = sum(num(r, "L2 Theoretical Sectors Global") for r in data), sum(num(r, "L2 Theoretical Sectors Global Ideal") for r in data)
    print(f"  global sectors/block {gs / nblk:.1f} (ideal {gi / nblk:.1f}), L1 tag requests/block {sum(num(r, 'L1 Tag Requests Global') for r in data) / nblk:.1f}")
    col = ci["Warp Stall Sampling (All Samples)"]
    tot = sum(num(r, "Warp Stall Sampling (All Samples)") for r in data) or 1
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    top = sorted(data, key=lambda r: -num(r, "Warp Stall Sampling (All Samples)"))[:6]
    for r in top:
        st = sorted(((num(r, h), h) for h in stall_cols), reverse=True)[0]
        print(f"  {100 * num(r, 'Warp Stall Sampling (All Samples)') / tot:5.1f} % of samples  {r[ci['Source']].strip()[:60]:60s} ({st[1]})")


if __name__ == "__main__":
    main()
