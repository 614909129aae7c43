#!/bin/bash
# round 2, GPU job S (1 GPU): TMA probe (which way of handing over a tensor map the part accepts), the tile-streaming kernel with its
# maps as a __grid_constant__ parameter: parity tests; compute-sanitizer on a small case if they fail; loop times at 3600x2400 if they pass
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== probe"; timeout 60 scripts/micro/tma_probe 2>&1 | tail -40
echo "== parity"
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tstream" 2>&1 | tail -12 > gpurun_out/r2_s_pytest.txt; cat gpurun_out/r2_s_pytest.txt
if grep -q "failed\|error" gpurun_out/r2_s_pytest.txt; then
  echo "== compute-sanitizer (gx3, 2 subcycles)"
  timeout 280 compute-sanitizer --tool memcheck --print-limit 6 python scripts/prof_step.py gx3 tstream exact 2 1 2>&1 | grep -v "^$" | head -60
else
  echo "== 3600x2400, 24 subcycles per loop"
  for k in stream tstream; do
    echo "-- $k (rows 12)"; timeout 200 python scripts/prof_step.py p1deg $k exact 24 4 2>&1 | tail -6
  done
  echo "-- tstream (rows 6)"; EVP_B200_TSTREAM_ROWS=6 timeout 200 python scripts/prof_step.py p1deg tstream exact 24 4 2>&1 | tail -6
  echo "== gx1 tstream vs fused (L2 resident), 240 subcycles"
  for k in fused tstream; do timeout 100 python scripts/prof_step.py gx1 $k exact 240 3 2>&1 | tail -3 | head -2; done
fi
} 2>&1 | tee gpurun_out/r2_s.txt
