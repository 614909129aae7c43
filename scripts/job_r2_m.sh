#!/bin/bash
# round 2, GPU job M (4 GPUs): persistent kernel with low-latency NVLink slots on a 2x2 processor grid (corner peers), configs[3] unchanged
mkdir -p gpurun_out
{
export EVP_B200_P2P_TIMEOUT_S=1
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29801 tests/mgpu_check.py gx3 25 29 41 persistent 2>&1 | grep "MGPU\|procs="  | cut -c1-200
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29802 tests/mgpu_check.py gx1 80 96 30 auto - elim 2>&1 | grep "MGPU\|procs=" | cut -c1-200
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29803 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2m_gx1_n4.json 2> gpurun_out/r2m_gx1_n4.err
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29804 bench.py --gpus 4 --workload tx1 --steps 10 --warmup 3 > gpurun_out/r2m_tx1_n4.json 2> gpurun_out/r2m_tx1_n4.err
for f in gx1_n4 tx1_n4; do python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2m_$f.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('$f', 'N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value %.3e'%d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity') and d['parity'].get('ok'), d['gpu_launches'], '|', d['config']['layout'][-160:])
except Exception as e:
    print('$f FAILED', e); print(open('gpurun_out/r2m_$f.err').read()[-1500:])
P
done
} 2>&1 | tee gpurun_out/r2_m.txt
