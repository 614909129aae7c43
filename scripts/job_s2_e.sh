#!/bin/bash
# session 2, job E: which half of the speculative form helps where (variants 16 none, 17 T loads, 18 cp.async U, 19 both)
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['e2e']['ms_per_step'])"; }
for v in 16 17 18 19; do
  echo "gx1 v$v: $(EVP_B200_FUSED_VARIANT=$v b --kernel fused)"
done
for v in 16 17 18 19; do
  echo "p1deg v$v: $(EVP_B200_FUSED_VARIANT=$v b --workload p1deg --steps 3)"
done
for v in 16 17 18 19; do
  echo "tx1 v$v: $(EVP_B200_FUSED_VARIANT=$v b --workload tx1)"
done
EVP_B200_FUSED_VARIANT=17 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused" 2>&1 | tail -1
EVP_B200_FUSED_VARIANT=18 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused" 2>&1 | tail -1
