#!/bin/bash
mkdir -p gpurun_out
for cfg in "64 1" "16 1" "32 1"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none -k regex:unpack_warp_kernel -s 2 -c 1 -f -o /tmp/prof_ud_$1 python tools/ncu_one.py undelta_pack $1 $2 $((32 - 13 - ($1 == 64 ? 0 : 0))) > gpurun_out/ncu_ud_$1.log 2>&1
  ncu -i /tmp/prof_ud_$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_undelta_u$1_w$2.csv 2>/dev/null
done
python - <<'PY'
import csv
keys=("gpu__time_duration.sum","sm__issue_active.avg.pct_of_peak_sustained_elapsed","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","smsp__inst_executed.sum","sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_membar_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_drain_per_issue_active.ratio","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","launch__grid_size")
for t in (64,16,32):
    rows=list(csv.reader(open(f"gpurun_out/ncu_raw_undelta_u{t}_w1.csv"))); d=dict(zip(rows[0],rows[2]))
    print("u%d"%t, {k.split("__")[-1][:40]: d.get(k) for k in keys})
PY
