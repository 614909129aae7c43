#!/bin/bash
# round 2, GPU job K (2 GPUs): evp_b200_step_resident on two ranks (staged velocity exchange after dyn_prep2, persistent kernel with the NVLink slots)
mkdir -p gpurun_out
{
export EVP_B200_P2P_TIMEOUT_S=1
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29801 tests/mgpu_check.py gx3 25 29 12 step 2>&1 | grep -v "^W1017\|^\*\*\*\|OMP_NUM" | grep "MGPU\|differs\|Error" | head -6 | cut -c1-250
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29802 tests/mgpu_check.py gx1 80 96 10 step 2>&1 | grep "MGPU\|differs" | head -5
} 2>&1 | tee gpurun_out/r2_k.txt
