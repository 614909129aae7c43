#!/bin/bash
# final round-1 evidence: launch list, ncu full captures (cold + warm) of the default kernel, C-grid kernels, bench lines
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
tail -c 600 gpurun_out/r1_bench_n1.json
python bench.py --steps 10 --warmup 3 --mode fast --no-cpu 2>/dev/null | tail -1 > gpurun_out/r1_bench_n1_fast.json
python bench.py --grid C --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r1_bench_cgrid.json
python bench.py --grid C --impl reference --steps 1 2>/dev/null | tail -1 > gpurun_out/r1_bench_cgrid_ref.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1_bench_ref.json
python bench.py --steps 4 --warmup 3 --workload p1deg --no-cpu 2>/dev/null | tail -1 > gpurun_out/r1_bench_p1deg_1gpu.json
EVP_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 40 --csv --log-file gpurun_out/r1_launches_fused_gx1.csv python scripts/prof_step.py gx1 fused exact 16 3 > /dev/null 2>&1
EVP_B200_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 20 -c 2 -o gpurun_out/r1_fused_cold -f python scripts/prof_step.py gx1 fused exact 16 3 > /dev/null 2>&1
EVP_B200_GRAPH=0 ncu --set full --cache-control none --clock-control none --import-source on -k regex:fused_kernel -s 20 -c 2 -o gpurun_out/r1_fused_warm -f python scripts/prof_step.py gx1 fused exact 16 3 > /dev/null 2>&1
ncu --set full --cache-control none --clock-control none -k regex:k[1-5]_ -s 50 -c 5 -o gpurun_out/r1_cgrid_warm -f python scripts/cgrid_time.py 40 > /dev/null 2>&1
ls -la gpurun_out | tail -12
