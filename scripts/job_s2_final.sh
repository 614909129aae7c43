#!/bin/bash
# session 2 evidence: bench lines of both arms, launch list of the bench command, ncu --set full (cold + warm) of the default fused kernel
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r1s2_bench_n1.json 2> gpurun_out/r1s2_bench_n1.err
tail -c 900 gpurun_out/r1s2_bench_n1.json; echo
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1s2_bench_ref.json
python bench.py --grid C --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r1s2_bench_cgrid.json
python bench.py --steps 10 --warmup 3 --mode fast --no-cpu 2>/dev/null | tail -1 > gpurun_out/r1s2_bench_n1_fast.json
python bench.py --steps 4 --warmup 3 --workload p1deg --no-cpu 2>/dev/null | tail -1 > gpurun_out/r1s2_bench_p1deg.json
python bench.py --steps 8 --warmup 3 --workload tx1 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r1s2_bench_tx1.json
python bench.py --steps 8 --warmup 3 --workload gx3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r1s2_bench_gx3.json
# launch list of the bench command itself (every kernel of 2 timed + 3 warm-up steps would be ~1500 launches: cap at 700, skip the upload)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1s2_launches_bench_gx1.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 300 -c 2 -o gpurun_out/r1s2_fused_cold -f python scripts/prof_step.py gx1 fused exact 240 2 > /dev/null 2>&1
timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:fused_kernel -s 300 -c 2 -o gpurun_out/r1s2_fused_warm -f python scripts/prof_step.py gx1 fused exact 240 2 > /dev/null 2>&1
timeout 600 ncu --set full --cache-control none --clock-control none -k regex:k[AB5]_ -s 60 -c 3 -o gpurun_out/r1s2_cgrid_warm -f python scripts/cgrid_time.py 40 > /dev/null 2>&1
ls -la gpurun_out | tail -14
