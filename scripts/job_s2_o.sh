#!/bin/bash
# session 2, job O: pairwise named barriers (v31 = 23 + pair, v32 = 19 + pair), sqrt_fast in stepu
for v in 31 32; do
  echo "parity v$v: $(EVP_B200_FUSED_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k 'fused or gx1_full or tripole or carry or split_api or boundary or max_blocks or edge_cases' 2>&1 | tail -1)"
done
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['e2e']['ms_per_step'])"; }
for v in 23 31 23 31; do echo "gx1 v$v: $(EVP_B200_FUSED_VARIANT=$v b)"; done
for v in 19 32; do echo "p1deg v$v: $(EVP_B200_FUSED_VARIANT=$v b --workload p1deg --steps 3)"; done
for v in 23 31; do echo "tx1 v$v: $(EVP_B200_FUSED_VARIANT=$v b --workload tx1)"; done
