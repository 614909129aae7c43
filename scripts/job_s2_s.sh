#!/bin/bash
# session 2, job S: cold paths out of line (div_ieee / sqrt_ieee / visc_tmp_general): parity + timing
echo "parity: $(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1)"
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['e2e']['ms_per_step'])"; }
echo "gx1: $(b)"; echo "gx1: $(b)"
echo "selftest 1: $(EVP_B200_P2P_SELFTEST=1 b)"
echo "selftest 2: $(EVP_B200_P2P_SELFTEST=2 b)"
echo "tx1: $(b --workload tx1)"
echo "p1deg: $(b --workload p1deg --steps 3)"
