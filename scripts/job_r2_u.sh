#!/bin/bash
# round 2, GPU job U (1 GPU): the TMA tile-streaming kernel with 16-byte aligned box starts (strip stride 30): probe of the alignment
# rule, parity tests, loop times at 3600x2400 against the launch-per-subcycle streaming form, ncu capture of one launch
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== probe"; timeout 60 scripts/micro/tma_probe 2>&1 | tail -40
echo "== parity"
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tstream" 2>&1 | tail -12 > gpurun_out/r2_u_pytest.txt; cat gpurun_out/r2_u_pytest.txt
if grep -q "failed\|error" gpurun_out/r2_u_pytest.txt; then
  echo "== compute-sanitizer (gx3, 2 subcycles)"
  timeout 280 compute-sanitizer --tool memcheck --print-limit 6 python scripts/prof_step.py gx3 tstream exact 2 1 2>&1 | grep -v "^$" | head -40
else
  echo "== 3600x2400, 24 subcycles per loop"
  for k in stream tstream; do
    echo "-- $k (rows 12)"; timeout 200 python scripts/prof_step.py p1deg $k exact 24 4 2>&1 | tail -6 | cut -c1-260
  done
  echo "-- tstream (rows 6)"; EVP_B200_TSTREAM_ROWS=6 timeout 200 python scripts/prof_step.py p1deg tstream exact 24 4 2>&1 | tail -6 | cut -c1-260
  echo "== gx1 tstream vs fused (L2 resident), 240 subcycles"
  for k in fused tstream; do timeout 100 python scripts/prof_step.py gx1 $k exact 240 3 2>&1 | tail -3 | head -2; done
  echo "== ncu, one launch at 3600x2400"
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:tstream_kernel -s 6 -c 1 -o gpurun_out/r2_tstream_p1deg -f python scripts/prof_step.py p1deg tstream exact 4 3 > gpurun_out/r2u_ncu.log 2>&1; tail -2 gpurun_out/r2u_ncu.log
fi
} 2>&1 | tee gpurun_out/r2_u.txt
