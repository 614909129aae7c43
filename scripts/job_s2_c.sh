#!/bin/bash
# session 2, job C: fp64 pipe microbenchmark + ncu full capture (with source) of the default fused kernel, warm L2
mkdir -p gpurun_out
scripts/micro/fp64_lat > gpurun_out/s2c_fp64_lat.txt 2>&1; cat gpurun_out/s2c_fp64_lat.txt
EVP_B200_GRAPH=0 timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:fused_kernel -s 20 -c 1 -o gpurun_out/s2c_fused_warm -f python scripts/prof_step.py gx1 fused exact 16 3 > gpurun_out/s2c_ncu.log 2>&1
tail -3 gpurun_out/s2c_ncu.log
ls -la gpurun_out/
