#!/bin/bash
# round 2, GPU job G (1 GPU): persistent kernel v3 -- parity, timing, per-warp cycle accounting, ncu capture
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "persistent" 2>&1 | tail -5
timeout 120 python scripts/prof_step.py gx1 persistent exact 240 4 2>&1 | tail -4
timeout 120 python scripts/prof_step.py gx1 fused exact 240 4 2>&1 | tail -2
EVP_B200_PERSIST_DEBUG=1 timeout 120 python scripts/prof_step.py gx1 persistent exact 240 2 2>&1 | tail -20
timeout 600 ncu --set full --import-source on --clock-control none -k regex:persist_kernel -s 1 -c 1 -o gpurun_out/r2_persist_v3 -f python scripts/prof_step.py gx1 persistent exact 240 3 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2_g.txt
