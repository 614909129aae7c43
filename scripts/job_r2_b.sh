#!/bin/bash
# round 2, GPU job B (one GPU): the pruned library -- full GPU suite, then the bench lines of every single-GPU workload.
mkdir -p gpurun_out
{
echo "== GPU suite"
timeout 1200 python -m pytest tests -m gpu -q -x -rs 2>&1 | tail -25
echo "== bench gx1 (default)"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_gx1.json 2> gpurun_out/r2b_bench_gx1.err; tail -c 600 gpurun_out/r2b_bench_gx1.err
echo "== bench tx1 / p1deg / gx3 / C grid"
timeout 600 python bench.py --workload tx1 --steps 8 --warmup 3 --no-cpu --no-pageable > gpurun_out/r2b_bench_tx1.json 2>gpurun_out/r2b_bench_tx1.err
timeout 900 python bench.py --workload p1deg --steps 3 --warmup 3 --no-cpu --no-pageable > gpurun_out/r2b_bench_p1deg.json 2>gpurun_out/r2b_bench_p1deg.err
timeout 600 python bench.py --workload gx3 --steps 8 --warmup 3 --no-cpu --no-pageable > gpurun_out/r2b_bench_gx3.json 2>gpurun_out/r2b_bench_gx3.err
timeout 600 python bench.py --grid C --steps 5 --warmup 3 > gpurun_out/r2b_bench_cgrid.json 2>gpurun_out/r2b_bench_cgrid.err
for f in gx1 tx1 p1deg gx3 cgrid; do python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r2b_bench_$f.json').read().strip().splitlines()[-1])
    print('$f', 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity',{}) and d['parity'].get('ok'), 'clocks', d.get('clocks'))
    if 'e2e_pageable' in d: print('   pageable', d['e2e_pageable']['pageable']['ms_per_step'], 'registered', d['e2e_pageable']['after_evp_b200_pin_host']['ms_per_step'])
except Exception as e:
    print('$f FAILED', e); print(open('gpurun_out/r2b_bench_$f.err').read()[-1500:])
P
done
} 2>&1 | tee gpurun_out/r2_b.txt
