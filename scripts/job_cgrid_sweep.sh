#!/bin/bash
python -m pytest tests/test_cgrid.py -m gpu -q 2>&1 | tail -4
python scripts/cgrid_time.py 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "exact_mode_bitwise_single_block and fused" 2>&1 | tail -2
EVP_B200_FUSED_VARIANT=20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(exact_mode_bitwise and fused) or (boundary_types) or (gx1_ndte240_every_kernel_exact and fused) or tripole_fold" 2>&1 | tail -3
for v in 0 20 11 12 13 14 15; do
  r=$(EVP_B200_FUSED_VARIANT=$v python bench.py --steps 8 --warmup 3 --kernel fused --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" 2>&1 | tail -1)
  echo "variant=$v ms_per_step,frac = $r"
done
EVP_B200_FUSED_VARIANT=20 python bench.py --steps 8 --warmup 3 --kernel fused --mode fast --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant 20 fast', d['ms_per_step'], d['roofline']['frac'])"
