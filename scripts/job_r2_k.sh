#!/bin/bash
# round 2, GPU job K (2 GPUs): multi-GPU parity tests of the persistent kernel and of the step-resident entry, bench at N = 2
mkdir -p gpurun_out
{
export EVP_B200_P2P_TIMEOUT_S=2
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_gpu and (persistent or auto or step)" 2>&1 | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29803 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2k_gx1_n2.json 2> gpurun_out/r2k_gx1_n2.err
python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2k_gx1_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('gx1_n2', 'ms/step', round(d['ms_per_step'],3), 'value %.3e'%d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), d['e2e']['how'][:40], 'resident_stress', round(d.get('e2e_resident_stress',{}).get('ms_per_step',0),3), 'parity', d['parity']['ok'], d['gpu_launches'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2k_gx1_n2.err').read()[-2500:])
P
} 2>&1 | tee gpurun_out/r2_k.txt
