#!/bin/bash
# session 2, job K (2 GPUs): multi-GPU parity (in-kernel NVLink halo + NCCL fallback + tripole + land-block elimination) and the 2-GPU weak-scaling bench
mkdir -p gpurun_out
nvidia-smi -L | head -4
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu" ) > gpurun_out/s2k_pytest.log 2>&1; tail -5 gpurun_out/s2k_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 tests/mgpu_check.py gx1 40 48 240 fused 2>&1 | grep -E "MGPU|differs|rror" | cut -c1-200 | head -4
timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/s2k_scale_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29732 bench.py --gpus 2 --steps 8 --warmup 3 2>gpurun_out/s2k_scale_n2.err | tail -1 > gpurun_out/s2k_scale_n2.json
EVP_B200_P2P_PDL=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 2 --steps 8 --warmup 3 2>/dev/null | tail -1 > gpurun_out/s2k_scale_n2_nopdl.json
for f in gpurun_out/s2k_scale_n1.json gpurun_out/s2k_scale_n2.json gpurun_out/s2k_scale_n2_nopdl.json; do
  python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value', '%.3e'%d['value'], 'e2e', '%.3e'%d['e2e']['value'], d['config']['layout'][-100:])
"
done
tail -3 gpurun_out/s2k_scale_n2.err
