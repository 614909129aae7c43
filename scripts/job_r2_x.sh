#!/bin/bash
# round 2, GPU job X (1 GPU): tile-streaming kernel with the box loads issued from uniform registers by an elected lane of warp 0
# (no per-load waterfall loop); the stress symmetrisation across the tripole fold on the device; loop times at 3600x2400
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== parity"
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tstream or symmetrised or refused_on_tripole or resident_stress" 2>&1 | tail -12 > gpurun_out/r2_x_pytest.txt; cat gpurun_out/r2_x_pytest.txt
echo "== 3600x2400, 24 subcycles per loop"
for r in 12 6; do
  echo "-- tstream rows $r"; EVP_B200_TSTREAM_ROWS=$r timeout 200 python scripts/prof_step.py p1deg tstream exact 24 5 2>&1 | tail -6 | cut -c1-300
done
} 2>&1 | tee gpurun_out/r2_x.txt
