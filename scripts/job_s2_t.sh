#!/bin/bash
# session 2, job T (4 GPUs): the PDL form of the in-kernel NVLink halo with 3 peers per rank: parity + weak-scaling bench line
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29720 tests/mgpu_check.py gx1 40 48 120 fused 2>&1 | grep -E "MGPU|differs|rror" | cut -c1-200 | head -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29721 tests/mgpu_check.py gx3 10 10 30 fused - elim 2>&1 | grep -E "MGPU|differs|rror" | cut -c1-200 | head -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29734 bench.py --gpus 4 --steps 6 --warmup 3 2>gpurun_out/s2t_scale_n4.err | tail -1 > gpurun_out/s2t_scale_n4.json
python -c "
import json
d=json.loads(open('gpurun_out/s2t_scale_n4.json').read().strip().splitlines()[-1])
print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value', '%.3e'%d['value'], 'e2e', '%.3e'%d['e2e']['value'], d['config']['layout'][-100:])
"
tail -2 gpurun_out/s2t_scale_n4.err
