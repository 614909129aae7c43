#!/bin/bash
# round 2, GPU job F (1 GPU): where the persistent kernel's time goes -- per-warp cycle accounting and one ncu --set full capture
mkdir -p gpurun_out
{
EVP_B200_PERSIST_DEBUG=1 timeout 120 python scripts/prof_step.py gx1 persistent exact 240 3 2>&1 | tail -45
timeout 600 ncu --set full --import-source on --clock-control none -k regex:persist_kernel -s 1 -c 1 -o gpurun_out/r2_persist_v2 -f python scripts/prof_step.py gx1 persistent exact 240 3 2>&1 | tail -5
} 2>&1 | tee gpurun_out/r2_f.txt
