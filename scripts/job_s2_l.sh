#!/bin/bash
# session 2, job L: fused C-grid form (kA = k1+k2, kB = k3+k4, k5): parity and timing against five kernels
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_cgrid.py -m gpu -q ) > gpurun_out/s2l_pytest.log 2>&1; tail -12 gpurun_out/s2l_pytest.log
echo "--- fused (3 kernels)"; timeout 300 python scripts/cgrid_time.py 600 2>&1 | tail -2
echo "--- five kernels"; EVP_B200_CGRID_FUSED=0 timeout 300 python scripts/cgrid_time.py 600 2>&1 | tail -1
