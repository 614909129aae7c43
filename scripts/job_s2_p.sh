#!/bin/bash
# session 2, job P (one GPU): cost of the in-kernel-halo kernel STRUCTURE without NVLink (EVP_B200_P2P_SELFTEST: 1 edge tiles first + counter,
# 2 natural order, 3 edge first no counter, 4 natural order + counter), with and without PDL between the subcycle kernels
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4))"; }
echo "plain fused: $(b)"
for st in 1 2 3 4; do
  for pdl in 0 1 2; do echo "selftest $st pdl $pdl: $(EVP_B200_P2P_SELFTEST=$st EVP_B200_P2P_PDL=$pdl b)"; done
done
echo "parity selftest 1 pdl 2: $(EVP_B200_P2P_SELFTEST=1 EVP_B200_P2P_PDL=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k 'exact_mode_bitwise_single_block and fused or gx1_ndte240_every_kernel_exact and fused' 2>&1 | tail -1)"
