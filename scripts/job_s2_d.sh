#!/bin/bash
# session 2, job D: speculative fused kernel (default) vs the non-speculative form (variant 16)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/s2d_pytest.log 2>&1; tail -3 gpurun_out/s2d_pytest.log
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['e2e']['ms_per_step'])"; }
echo "fused spec (default): $(b --kernel fused)"
echo "fused spec nopdl: $(EVP_B200_PDL=0 b --kernel fused)"
echo "fused v16 (old): $(EVP_B200_FUSED_VARIANT=16 b --kernel fused)"
echo "fused spec fast: $(b --kernel fused --mode fast)"
echo "p1deg spec: $(b --workload p1deg --steps 3)"
echo "p1deg v16: $(EVP_B200_FUSED_VARIANT=16 b --workload p1deg --steps 3)"
echo "gx3 spec: $(b --workload gx3)"; echo "gx3 v16: $(EVP_B200_FUSED_VARIANT=16 b --workload gx3)"
echo "tx1 spec: $(b --workload tx1)"; echo "tx1 v16: $(EVP_B200_FUSED_VARIANT=16 b --workload tx1)"
