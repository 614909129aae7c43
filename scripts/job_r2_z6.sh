#!/bin/bash
# round 2, the last GPU seconds (1 GPU): the default bench command once more on the very last tree
mkdir -p gpurun_out
{
timeout 100 python bench.py --no-pageable > gpurun_out/r2z6_bench_gx1.json 2> gpurun_out/r2z6_bench_gx1.err; tail -2 gpurun_out/r2z6_bench_gx1.err | cut -c1-300
python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2z6_bench_gx1.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('gx1 value %.4e'%d['value'], 'ms', round(d['ms_per_step'],3), 'kernel', d['config']['kernel'][:40], 'e2e', round(d['e2e'].get('ms_per_step',0),3), 'parity', (d.get('parity') or {}).get('ok'), 'launches', d['gpu_launches'], 'cpu', d.get('cpu_baseline',{}).get('value'))
except Exception as e:
    print('FAILED', e)
P
} 2>&1 | tee gpurun_out/r2_z6.txt
