#!/bin/bash
# session 2, job H: early PDL trigger, CTA shapes with the interleaved div/sqrt form
echo "parity (trigger on): $(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1)"
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['e2e']['ms_per_step'])"; }
echo "gx1 v23 trigger: $(b)"
echo "gx1 v23 no trigger: $(EVP_B200_PDL_TRIGGER=0 b)"
echo "gx1 v23 no pdl: $(EVP_B200_PDL=0 b)"
for v in 17 21 19 22 24 25 26 27 28 29; do echo "gx1 v$v trigger: $(EVP_B200_FUSED_VARIANT=$v b)"; done
echo "p1deg v19 trigger: $(b --workload p1deg --steps 3)"
echo "p1deg v19 no trigger: $(EVP_B200_PDL_TRIGGER=0 b --workload p1deg --steps 3)"
echo "tx1 trigger: $(b --workload tx1)"; echo "tx1 no trigger: $(EVP_B200_PDL_TRIGGER=0 b --workload tx1)"
echo "gx3 trigger: $(b --workload gx3)"; echo "gx3 no trigger: $(EVP_B200_PDL_TRIGGER=0 b --workload gx3)"
