#!/bin/bash
# round 2, GPU job I (1 GPU): ncu capture of the persistent kernel (v4)
mkdir -p gpurun_out
{
timeout 600 ncu --set full --import-source on --clock-control none -k regex:persist_kernel -s 1 -c 1 -o gpurun_out/r2_persist_v4 -f python scripts/prof_step.py gx1 persistent exact 240 3 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2_i.txt
