#!/bin/bash
# round-2 call X (1 GPU): where do the shared-memory bank conflicts of the u64 original-order chains at W = 1 come from?
# ncu --set full with the source page (per-instruction shared wavefronts) for transpose_delta_pack and undelta_pack_untranspose
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
cap() {  # name, kernel regex, op, T, W
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o /tmp/prof_$1 python tools/ncu_one.py $3 $4 $5 19 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/ncu_source_$1.csv 2>/dev/null
}
cap tdp_u64_w1 pack_warp_kernel transpose_delta_pack 64 1
cap udo_u64_w1 unpack_warp_kernel undelta_pack_untranspose 64 1
python tools/ncu_digest.py tdp_u64_w1 udo_u64_w1 > gpurun_out/ncu_digest_x.md; cat gpurun_out/ncu_digest_x.md
python - <<'PY'
import csv
for name in ("tdp_u64_w1", "udo_u64_w1"):
    rows = list(csv.reader(open(f"gpurun_out/ncu_source_{name}.csv")))
    hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
    print("==", name)
    for r in rows[2:]:
        try:
            w, ideal, ex = int(r[ci["L1 Wavefronts Shared"]]), int(r[ci["L1 Wavefronts Shared Ideal"]]), int(r[ci["L1 Wavefronts Shared Excessive"]])
        except ValueError:
            continue
        if w:
            print(f"{r[ci['Source']].strip()[:70]:70s} executed {r[ci['Instructions Executed']]:>9s} wavefronts {w:>10d} ideal {ideal:>10d} excess {ex:>10d}")
PY
