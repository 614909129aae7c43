#!/bin/bash
# round 2, GPU job H (1 GPU): persistent kernel -- timing + timeline maps
mkdir -p gpurun_out
{
timeout 120 python scripts/prof_step.py gx1 persistent exact 240 3 2>&1 | tail -2
timeout 120 python scripts/prof_step.py gx1 fused exact 240 3 2>&1 | tail -1
EVP_B200_PERSIST_DEBUG=2 timeout 120 python scripts/prof_step.py gx1 persistent exact 240 2 2>&1 | tail -95
} 2>&1 | tee gpurun_out/r2_h.txt
