#!/bin/bash
# round 2, first GPU job (one GPU): the candidates written at the end of round 1 without a GPU.
#   1. parity of the two-lanes-per-cell kernel variants and of evp_b200_dyn_finish (host-emulated so far: tests/test_emu_bgrid.py)
#   2. gx1 bench line per variant (ms per step, roofline fraction) next to the default
#   3. the in-kernel-halo kernel with its tile table in constant memory, in the one-GPU no-peer self test
# Output: gpurun_out/r2_candidates.txt.   /usr/local/graft/bin/gpurun --timeout 1500 -- bash scripts/job_r2_candidates.sh
mkdir -p gpurun_out
{
echo "== parity of variants 40-53"
EVP_B200_TEST_CANDIDATES=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_lane or dyn_finish or derived_geometry or tripole_fold_as_one or pinning" 2>&1 | tail -15
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4))"; }
EVP_B200_TEST_CANDIDATES=1 timeout 600 python -m pytest tests/test_cgrid.py -m gpu -q -k programmatic 2>&1 | tail -3
echo "== gx1 CD grid (ndte=600): four kernels per subcycle, plain vs programmatic dependent launch"
echo "cd plain: $(timeout 300 python scripts/cdgrid_time.py 2>&1 | tail -1)"
echo "cd pdl:   $(EVP_B200_CDGRID_PDL=1 timeout 300 python scripts/cdgrid_time.py 2>&1 | tail -1)"
echo "== gx1 C grid (ndte=600): default vs programmatic dependent launch"
for sh in 0 5 16 17 18 19; do echo "cgrid shape $sh: $(EVP_B200_CGRID_SHAPE=$sh b --grid C)"; done
echo "== gx1, ms per step and roofline fraction"
echo "default: $(b)"
for v in 40 41 42 43 44 45 46 47 48 49 50 51 52 53 54 55 56; do echo "variant $v: $(EVP_B200_FUSED_VARIANT=$v b)"; done
echo "== derived geometry (two metric arrays instead of seven): gx1 variant 63 vs 23, 3600x2400 variant 59 vs 19"
echo "gx1 v63: $(EVP_B200_FUSED_VARIANT=63 b)"
echo "p1deg v19 (default): $(b --workload p1deg --steps 3 --warmup 2)"
echo "p1deg v59: $(EVP_B200_FUSED_VARIANT=59 b --workload p1deg --steps 3 --warmup 2)"
echo "== tx1 tripole on one GPU: fold as pack + apply (default) vs one kernel vs one kernel in the PDL chain"
for hf in 0 1 2; do echo "tx1 halo_fused $hf: $(EVP_B200_HALO_FUSED=$hf b --workload tx1)"; done
echo "== gx1 fast mode"
echo "default fast: $(b --mode fast)"
for v in 40 43 50; do echo "variant $v fast: $(EVP_B200_FUSED_VARIANT=$v b --mode fast)"; done
echo "== in-kernel-halo kernel, no-peer self test: tile table in global vs constant memory"
for st in 1 3; do for ct in 0 1; do echo "selftest $st const_tiles $ct: $(EVP_B200_P2P_SELFTEST=$st EVP_B200_P2P_CONST_TILES=$ct b)"; done; done
echo "selftest 1, two-lane kernel (variant 40): $(EVP_B200_P2P_SELFTEST=1 EVP_B200_FUSED_VARIANT=40 b)"
echo "parity selftest 1 + two-lane kernel: $(EVP_B200_P2P_SELFTEST=1 EVP_B200_FUSED_VARIANT=40 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k 'exact_mode_bitwise and fused or gx1_ndte240_every_kernel_exact and fused or boundary_types' 2>&1 | tail -1)"
echo "parity selftest 1 + constant tiles: $(EVP_B200_P2P_SELFTEST=1 EVP_B200_P2P_CONST_TILES=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k 'exact_mode_bitwise and fused or gx1_ndte240_every_kernel_exact and fused or boundary_types' 2>&1 | tail -1)"
} 2>&1 | tee gpurun_out/r2_candidates.txt
