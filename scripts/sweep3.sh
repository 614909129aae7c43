#!/bin/bash
for v in 0 11 12 13 14 15; do
  r=$(EVP_B200_FUSED_VARIANT=$v python bench.py --steps 8 --warmup 3 --kernel fused --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" 2>&1 | tail -1)
  echo "variant=$v ms_per_step,frac = $r"
done
