#!/bin/bash
# round 2, GPU job H (1 GPU): persistent kernel -- parity + timing
mkdir -p gpurun_out
{
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "persistent or step_prep or repeated" 2>&1 | tail -2
timeout 120 python scripts/prof_step.py gx1 persistent exact 240 4 2>&1 | tail -3
timeout 120 python scripts/prof_step.py gx3 persistent exact 120 3 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2_h.txt
