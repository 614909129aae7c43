#!/bin/bash
# 8-GPU box: parity on 4 and 8 ranks, then the weak-scaling bench at N=1,2,4,8 (NVLink stores) and N=8 with NCCL
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n tests/mgpu_check.py gx3 25 29 60 fused 2>&1 | grep -E "MGPU|differs|rror|procs" | cut -c1-160 | head -4
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 tests/mgpu_check.py gx1 40 48 240 fused 2>&1 | grep -E "MGPU|differs|rror" | cut -c1-160 | head -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29721 tests/mgpu_check.py tiny 6 5 16 fused tripole 2>&1 | grep -E "MGPU|differs|rror" | cut -c1-160 | head -4
python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/scale_n1.json 2>gpurun_out/scale_n1.err
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2973$n bench.py --gpus $n --steps 8 --warmup 3 2>gpurun_out/scale_n$n.err | tail -1 > gpurun_out/scale_n$n.json
done
EVP_B200_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29740 bench.py --gpus 8 --steps 8 --warmup 3 2>/dev/null | tail -1 > gpurun_out/scale_n8_nccl.json
for f in gpurun_out/scale_n1.json gpurun_out/scale_n2.json gpurun_out/scale_n4.json gpurun_out/scale_n8.json gpurun_out/scale_n8_nccl.json; do
  python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value', '%.3e'%d['value'], 'e2e', '%.3e'%d['e2e']['value'])
"
done
