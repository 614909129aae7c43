#!/bin/bash
# round 2, GPU job J (1 GPU): whole GPU suite with the persistent kernel as AUTO's choice, default bench line, smoke
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/r2j_bench_gx1.json 2> gpurun_out/r2j_bench_gx1.err; tail -c 1500 gpurun_out/r2j_bench_gx1.json
} 2>&1 | tee gpurun_out/r2_j.txt
