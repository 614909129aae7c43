#!/bin/bash
# session 2, job R: ncu comparison plain fused kernel vs the in-kernel-halo template without peers (self test 2 and 1)
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__cycles_active.avg,sm__cycles_elapsed.avg,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio,sm__warps_active.avg.per_cycle_active,lts__t_sectors.sum,launch__grid_size
for st in 0 2 1; do
  if [ $st = 0 ]; then envs=""; else envs="EVP_B200_P2P_SELFTEST=$st"; fi
  env $envs timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:fused_kernel -s 300 -c 2 --csv --log-file gpurun_out/s2r_$st.csv python scripts/prof_step.py gx1 fused exact 240 2 > /dev/null 2>&1
done
python - <<'PY'
import csv
for st in (0,2,1):
    rows=[r for r in csv.reader(open(f"gpurun_out/s2r_{st}.csv")) if len(r)>14 and r[0].isdigit()]
    m={}
    for r in rows:
        if r[0]=="0": m[r[12]]=r[14]
    print("selftest",st, rows[0][4][:70] if rows else None)
    for k,v in m.items(): print("   ",k.replace("smsp__average_warps_issue_stalled_","stall_").replace("_per_issue_active.ratio",""),v)
PY
