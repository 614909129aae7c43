#!/bin/bash
# round 2, GPU job O (1 GPU): streaming form of the fused kernel with the bulk L2 prefetch at several distances, 3600x2400
mkdir -p gpurun_out
{
for pf in 0 148 296 444 888 1776; do
  echo "prefetch distance $pf tiles:"; EVP_B200_PREFETCH_TILES=$pf timeout 200 python scripts/prof_step.py p1deg stream exact 24 3 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/r2_o.txt
