#!/bin/bash
# round 2, last GPU minutes (2 GPUs): tx1 tripole cut in two (strong layout), the end-to-end leg with the stresses resident and
# symmetrised across the fold between the two ranks of the top row
mkdir -p gpurun_out
{
nvidia-smi -L
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload tx1 --steps 10 --warmup 3 --no-cpu --no-pageable > gpurun_out/r2z_bench_tx1_n2.json 2> gpurun_out/r2z_bench_tx1_n2.err; tail -3 gpurun_out/r2z_bench_tx1_n2.err | cut -c1-300
python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2z_bench_tx1_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('tx1 n2 value %.4e'%d['value'], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e'].get('ms_per_step',0),3), d['e2e']['how'][:80], 'full copy', round(d['e2e_full_copy']['ms_per_step'],3), 'parity', (d.get('parity') or {}).get('ok'), d['config']['workload'])
except Exception as e:
    print('FAILED', e)
P
} 2>&1 | tee gpurun_out/r2_z5.txt
