#!/bin/bash
# session 2, job F: interleaved IEEE division / square root (variants 21 = 17+IL, 22 = 19+IL, 23 = 16+IL): parity then timing
for v in 21 22; do
  echo "parity v$v: $(EVP_B200_FUSED_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k 'fused or gx1_full or tripole or carry or split_api or boundary or max_blocks or edge_cases' 2>&1 | tail -1)"
done
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['e2e']['ms_per_step'])"; }
for v in 17 21 23 22; do echo "gx1 v$v: $(EVP_B200_FUSED_VARIANT=$v b --kernel fused)"; done
for v in 19 22; do echo "p1deg v$v: $(EVP_B200_FUSED_VARIANT=$v b --workload p1deg --steps 3)"; done
for v in 17 21; do echo "tx1 v$v: $(EVP_B200_FUSED_VARIANT=$v b --workload tx1)"; done
for v in 17 21; do echo "gx3 v$v: $(EVP_B200_FUSED_VARIANT=$v b --workload gx3)"; done
