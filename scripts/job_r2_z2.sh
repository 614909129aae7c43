#!/bin/bash
# round 2, last GPU job (2 GPUs): the N = 2 bench line exactly as the driver launches it, on the final tree
mkdir -p gpurun_out
{
nvidia-smi -L
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2z_bench_gx1_n2.json 2> gpurun_out/r2z_bench_gx1_n2.err; tail -3 gpurun_out/r2z_bench_gx1_n2.err | cut -c1-300
python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2z_bench_gx1_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('n2 value %.4e'%d['value'], 'ms', round(d['ms_per_step'],3), 'kernel', d['config']['kernel'], 'e2e', round(d['e2e'].get('ms_per_step',0),3), 'parity', (d.get('parity') or {}).get('ok'), 'clocks', d.get('clocks'))
except Exception as e:
    print('FAILED', e)
P
} 2>&1 | tee gpurun_out/r2_z2.txt
