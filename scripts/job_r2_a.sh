#!/bin/bash
# round 2, GPU job A (one GPU): every opt-in candidate test of round 1 on hardware, then bench lines of the candidates worth keeping.
mkdir -p gpurun_out
{
echo "== full GPU suite with candidates"
EVP_B200_TEST_CANDIDATES=1 timeout 1200 python -m pytest tests -m gpu -q -rfEs 2>&1 | tail -60
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['e2e']['ms_per_step'])"; }
echo "== gx1, ms per step and roofline fraction"
echo "default: $(b)"
for v in 40 43 47 50 54 55 56 63; do echo "variant $v: $(EVP_B200_FUSED_VARIANT=$v b)"; done
echo "== p1deg"
echo "p1deg v19 (default): $(b --workload p1deg --steps 3 --warmup 2)"
echo "p1deg v59: $(EVP_B200_FUSED_VARIANT=59 b --workload p1deg --steps 3 --warmup 2)"
echo "== tx1 tripole on one GPU"
for hf in 0 1 2; do echo "tx1 halo_fused $hf: $(EVP_B200_HALO_FUSED=$hf b --workload tx1)"; done
echo "== C grid"
for sh in 0 5 16 17; do echo "cgrid shape $sh: $(EVP_B200_CGRID_SHAPE=$sh b --grid C)"; done
echo "== selftest"
for st in 1 3; do for ct in 0 1; do echo "selftest $st const_tiles $ct: $(EVP_B200_P2P_SELFTEST=$st EVP_B200_P2P_CONST_TILES=$ct b)"; done; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc; lscpu | grep "Model name"
} 2>&1 | tee gpurun_out/r2_a.txt
