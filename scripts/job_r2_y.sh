#!/bin/bash
# round 2, GPU job Y (1 GPU): A/B of the tile-streaming kernel's variants on one box in one process
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 400 python scripts/ab_tstream.py p1deg 24 3 2>&1 | tail -40
} 2>&1 | tee gpurun_out/r2_y.txt
