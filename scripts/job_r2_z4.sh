#!/bin/bash
# round 2, closing GPU job (1 GPU): the GPU suite on the final tree once more (after the multi-rank stress symmetrisation went in) and
# the tx1 bench line, whose end-to-end leg now keeps the stresses on the device (tripole grid, symmetrised there)
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py --workload tx1 --steps 10 --warmup 3 --no-cpu --no-pageable > gpurun_out/r2z_bench_tx1.json 2> gpurun_out/r2z_bench_tx1.err; tail -2 gpurun_out/r2z_bench_tx1.err
python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2z_bench_tx1.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('tx1 value %.4e'%d['value'], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e'].get('ms_per_step',0),3), d['e2e']['how'][:90], 'full copy', round(d['e2e_full_copy']['ms_per_step'],3), 'parity', (d.get('parity') or {}).get('ok'))
except Exception as e:
    print('FAILED', e)
P
} 2>&1 | tee gpurun_out/r2_z4.txt
