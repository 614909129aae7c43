#!/bin/bash
# round 2, GPU job Q (4 GPUs): every multi-GPU parity test on a 2x2 processor grid (nothing skipped for lack of GPUs); log kept in profiles/
mkdir -p gpurun_out
{
nvidia-smi -L
export EVP_B200_P2P_TIMEOUT_S=2
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -v -rs -k "multi_gpu" 2>&1 | grep -v "^$" | tail -40
} 2>&1 | tee gpurun_out/r2_pytest_multi_gpu_4xB200.log
