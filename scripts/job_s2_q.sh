#!/bin/bash
# session 2, job Q (one GPU): packed tile table (no integer division), PDL, in the no-peer self test of the in-kernel-halo kernel
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4))"; }
echo "plain: $(b)"
for st in 1 2 3 4; do for pdl in 0 2; do echo "selftest $st pdl $pdl: $(EVP_B200_P2P_SELFTEST=$st EVP_B200_P2P_PDL=$pdl b)"; done; done
echo "parity selftest 1: $(EVP_B200_P2P_SELFTEST=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k 'exact_mode_bitwise and fused or gx1_ndte240_every_kernel_exact and fused or boundary_types' 2>&1 | tail -1)"
