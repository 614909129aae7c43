#!/bin/bash
# session 2, job B: instruction diet (32-bit indices, TbU shortcut, diagnostics only in the last subcycle, uvel_init elision)
# and the strip kernel (variant 30), parity + timing
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/s2b_pytest.log 2>&1; tail -3 gpurun_out/s2b_pytest.log
for m in 1 2 3 5; do
  ( EVP_B200_FUSED_VARIANT=30 EVP_B200_STRIP_M=$m timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or gx1_full or tripole or carry or split_api or boundary or max_blocks" ) > gpurun_out/s2b_pytest_strip_m$m.log 2>&1; echo "strip m=$m: $(tail -1 gpurun_out/s2b_pytest_strip_m$m.log)"
done
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['e2e']['ms_per_step'])"; }
echo "fused default: $(b --kernel fused)"
for m in 1 2 3 4; do echo "strip m=$m: $(EVP_B200_FUSED_VARIANT=30 EVP_B200_STRIP_M=$m b --kernel fused)"; done
echo "strip auto: $(EVP_B200_FUSED_VARIANT=30 b --kernel fused)"
echo "strip auto nopdl: $(EVP_B200_FUSED_VARIANT=30 EVP_B200_PDL=0 b --kernel fused)"
echo "strip auto fast: $(EVP_B200_FUSED_VARIANT=30 b --kernel fused --mode fast)"
echo "queue: $(b --kernel queue)"
echo "p1deg fused: $(b --workload p1deg --steps 3)"
for m in 2 4 8; do echo "p1deg strip m=$m: $(EVP_B200_FUSED_VARIANT=30 EVP_B200_STRIP_M=$m b --workload p1deg --steps 3)"; done
echo "gx3 fused: $(b --workload gx3)"; echo "gx3 strip: $(EVP_B200_FUSED_VARIANT=30 b --workload gx3)"
echo "tx1 fused: $(b --workload tx1)"; echo "tx1 strip: $(EVP_B200_FUSED_VARIANT=30 b --workload tx1)"
