#!/bin/bash
# round 2: ncu --set full (warm L2, source-level) of the two-lanes-per-cell kernel next to the default, gx1 exact -- run after
# scripts/job_r2_candidates.sh has shown which variant is the fastest (default here: 54 = specialised roles; pass another as $1).
#   /usr/local/graft/bin/gpurun --timeout 900 -- bash scripts/job_r2_profile_lane2.sh 54
V=${1:-54}
mkdir -p gpurun_out
EVP_B200_GRAPH=0 EVP_B200_FUSED_VARIANT=$V ncu --set full --cache-control none --clock-control none --import-source on -k regex:fused2_kernel -s 20 -c 2 \
  -o gpurun_out/r2_lane2_v${V}_warm -f python scripts/prof_step.py gx1 fused exact 16 3 > /dev/null 2>&1
EVP_B200_GRAPH=0 ncu --set full --cache-control none --clock-control none --import-source on -k regex:fused_kernel -s 20 -c 2 \
  -o gpurun_out/r2_default_warm -f python scripts/prof_step.py gx1 fused exact 16 3 > /dev/null 2>&1
for f in r2_lane2_v${V}_warm r2_default_warm; do
  python scripts/ncu_summary.py gpurun_out/$f.ncu-rep gpurun_out/$f.txt "round 2, gx1 exact, --cache-control none" 2>&1 | tail -3; tail -25 gpurun_out/$f.txt
done
ls -la gpurun_out | tail -6
