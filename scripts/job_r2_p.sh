#!/bin/bash
# round 2, GPU job P (8 GPUs): weak scaling at gx1 per GPU with the persistent kernel + low-latency NVLink slots on the 4x2 processor grid
mkdir -p gpurun_out
{
export EVP_B200_P2P_TIMEOUT_S=2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2p_gx1_n8.json 2> gpurun_out/r2p_gx1_n8.err
python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2p_gx1_n8.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('gx1_n8', 'N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value %.3e'%d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity') and d['parity'].get('ok'), d['gpu_launches'], d['clocks'], '|', d['config']['layout'][-150:])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2p_gx1_n8.err').read()[-2500:])
P
} 2>&1 | tee gpurun_out/r2_p.txt
