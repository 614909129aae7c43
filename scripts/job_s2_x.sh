#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_cgrid.py -m gpu -q -k "deformations or cgrid_exact_bitwise or cdgrid_exact_bitwise" 2>&1 | tail -4 | cut -c1-200
