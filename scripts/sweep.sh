#!/bin/bash
# usage: scripts/sweep.sh  -- fused-kernel variants x PDL, gx1 exact
for pdl in 1 0; do for v in 0 1 2 3 4 5 6; do
  r=$(EVP_B200_PDL=$pdl EVP_B200_FUSED_VARIANT=$v python bench.py --steps 5 --warmup 3 --kernel fused --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" 2>&1 | tail -1)
  echo "pdl=$pdl variant=$v ms_per_step,frac = $r"
done; done
