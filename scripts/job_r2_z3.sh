#!/bin/bash
# round 2, GPU job (2 GPUs): stresses resident on a tripole grid whose top row is spread over two ranks -- the symmetrisation across
# the fold with the top-row segments exchanged between the ranks (NCCL), two steps, against the oracle
mkdir -p gpurun_out
{
nvidia-smi -L
export EVP_B200_P2P_TIMEOUT_S=2
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -v -rs -k "multi_gpu and resident-stresses and nvlink" 2>&1 | grep -v "^$" | tail -30 | cut -c1-400
} 2>&1 | tee gpurun_out/r2_z3.txt
