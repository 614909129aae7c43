#!/bin/bash
# round 2, GPU job D (4 GPUs of one box): configs[3] (tx1 360x240 tripole, strong 2x2) and weak scaling gx1 at 2 and 4 GPUs, every line
# with the parity pre-check.   /usr/local/graft/bin/gpurun --gpus 4 --timeout 900 -- bash scripts/job_r2_d.sh
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
{
nvidia-smi -L
run 4 29801 bench.py --gpus 4 --workload tx1 --steps 10 --warmup 3 > gpurun_out/r2d_tx1_n4.json 2> gpurun_out/r2d_tx1_n4.err
EVP_B200_P2P=0 run 4 29802 bench.py --gpus 4 --workload tx1 --steps 10 --warmup 3 > gpurun_out/r2d_tx1_n4_nccl.json 2> gpurun_out/r2d_tx1_n4_nccl.err
run 2 29803 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_gx1_n2.json 2> gpurun_out/r2d_gx1_n2.err
run 4 29804 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2d_gx1_n4.json 2> gpurun_out/r2d_gx1_n4.err
EVP_B200_P2P_DEBUG=1 run 4 29805 bench.py --gpus 4 --steps 3 --warmup 3 --no-parity > /dev/null 2> gpurun_out/r2d_gx1_n4_debug.err
grep "p2p rank 0" gpurun_out/r2d_gx1_n4_debug.err | tail -3
for f in tx1_n4 tx1_n4_nccl gx1_n2 gx1_n4; do python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2d_$f.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('$f', 'N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value %.3e'%d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity') and d['parity'].get('ok'), '|', d['config']['layout'][-110:])
except Exception as e:
    print('$f FAILED', e); print(open('gpurun_out/r2d_$f.err').read()[-1500:])
P
done
} 2>&1 | tee gpurun_out/r2_d.txt
