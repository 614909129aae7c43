#!/bin/bash
# round 2, GPU job R (1 GPU): first hardware run of the TMA tile-streaming kernel (KERNEL_TSTREAM): parity tests, then loop times at
# 3600x2400 against the launch-per-subcycle streaming form, both instantiations
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tstream or derived_geometry" 2>&1 | tail -15
echo "== 3600x2400, 24 subcycles per loop"
for k in stream tstream; do
  echo "-- $k (rows 12)"; timeout 200 python scripts/prof_step.py p1deg $k exact 24 4 2>&1 | tail -6
done
echo "-- tstream (rows 6)"; EVP_B200_TSTREAM_ROWS=6 timeout 200 python scripts/prof_step.py p1deg tstream exact 24 4 2>&1 | tail -6
} 2>&1 | tee gpurun_out/r2_r.txt
