#!/bin/bash
# round 2, GPU job W (1 GPU): tile-streaming kernel with a producer warp for the box loads: parity, loop times at 3600x2400
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== parity"
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tstream" 2>&1 | tail -12 > gpurun_out/r2_w_pytest.txt; cat gpurun_out/r2_w_pytest.txt
if ! grep -q "failed\|error" gpurun_out/r2_w_pytest.txt; then
  echo "== 3600x2400, 24 subcycles per loop"
  for r in 12 6; do
    echo "-- tstream rows $r"; EVP_B200_TSTREAM_ROWS=$r timeout 200 python scripts/prof_step.py p1deg tstream exact 24 5 2>&1 | tail -6 | cut -c1-300
  done
fi
} 2>&1 | tee gpurun_out/r2_w.txt
