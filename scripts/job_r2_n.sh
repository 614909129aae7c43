#!/bin/bash
# round 2, GPU job N (1 GPU): default bench line with the step-resident end-to-end leg
mkdir -p gpurun_out
{
timeout 300 python bench.py > gpurun_out/r2n_bench_gx1.json 2> gpurun_out/r2n_bench_gx1.err; tail -3 gpurun_out/r2n_bench_gx1.err
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r2n_bench_gx1.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value %.4e'%d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], 'resident_stress', d['e2e_resident_stress']['ms_per_step'], 'full', d['e2e_full_copy']['ms_per_step'], 'pageable', d['e2e_pageable']['pageable']['ms_per_step'], d['e2e_pageable']['after_evp_b200_pin_host']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], d['parity']['ok'], d['clocks'])
P
} 2>&1 | tee gpurun_out/r2_n.txt
