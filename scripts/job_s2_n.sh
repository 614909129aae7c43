#!/bin/bash
# session 2, job N: C-grid fused form, tile shapes / residency
for sh in 0 1 2 3 4; do
  echo "shape $sh: $(EVP_B200_CGRID_SHAPE=$sh timeout 300 python scripts/cgrid_time.py 600 2>&1 | tail -1 | cut -c1-80)"
done
echo "parity shape 1: $(EVP_B200_CGRID_SHAPE=1 timeout 300 python -m pytest tests/test_cgrid.py -m gpu -q -k 'exact_bitwise or gx1' 2>&1 | tail -1)"
echo "parity shape 2: $(EVP_B200_CGRID_SHAPE=2 timeout 300 python -m pytest tests/test_cgrid.py -m gpu -q -k 'exact_bitwise or gx1' 2>&1 | tail -1)"
