#!/bin/bash
# round-2 call AB (1 GPU): measurement evidence for the final library — ncu launch list of the bench command, and
# ncu --set full captures of the headline kernel (W = 1, 16, 32) + filter / select W = 8 for DRAM traffic per launch
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/ncu_raw_*.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu --no-other > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list exit $?"
cap() {  # name, kernel regex, op, T, W
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o /tmp/prof_$1 python tools/ncu_one.py $3 $4 $5 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
}
cap unpack_u32_w1 unpack_warp_kernel unpack 32 1
cap unpack_u32_w16 unpack_warp_kernel unpack 32 16
cap unpack_u32_w32 unpack_warp_kernel unpack 32 32
cap filter_u32_w8 filter_warp_kernel unpack_filter 32 8
cap select_u32_w8 select_warp_kernel unpack_select 32 8
python tools/ncu_digest.py unpack_u32_w1 unpack_u32_w16 unpack_u32_w32 filter_u32_w8 select_u32_w8 > gpurun_out/ncu_digest_ab.md; cat gpurun_out/ncu_digest_ab.md
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_r02_final.csv")))
i = next(k for k, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[i]; ci = {h: k for k, h in enumerate(hdr)}
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[i + 1:]:
    if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ci["Metric Value"]].replace(",", "")); u = r[ci["Metric Unit"]]
    us = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    k = r[ci["Kernel Name"]].split("(")[0][:60]
    tot[k][0] += 1; tot[k][1] += us
s = sum(v[1] for v in tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} launches {v[0]:4d} total {v[1]:10.1f} us share {100 * v[1] / s:5.1f} %")
PY
