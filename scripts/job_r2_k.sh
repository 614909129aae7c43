#!/bin/bash
# round 2, GPU job K (2 GPUs): the persistent kernel with neighbour GPUs (low-latency slots) -- one parity check, weak-scaling bench at N=2
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
{
export EVP_B200_P2P_TIMEOUT_S=1
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29801 tests/mgpu_check.py gx3 25 29 41 persistent 2>&1 | grep -v "^W1017\|^\*\*\*\|OMP_NUM" | tail -6
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29803 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2k_gx1_n2.json 2> gpurun_out/r2k_gx1_n2.err
for f in gx1_n2; do python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2k_$f.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('$f', 'N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value %.3e'%d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity') and d['parity'].get('ok'), d['gpu_launches'], '|', d['config']['layout'][-160:])
except Exception as e:
    print('$f FAILED', e); print(open('gpurun_out/r2k_$f.err').read()[-1500:])
P
done
} 2>&1 | tee gpurun_out/r2_k.txt
