#!/bin/bash
# round 2, GPU job K (2 GPUs): multi-GPU parity tests of the persistent kernel and of the step-resident entry (odd ndte: the carried
# state ends on the other ping-pong copy, the neighbours' copies must follow)
mkdir -p gpurun_out
{
export EVP_B200_P2P_TIMEOUT_S=2
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_gpu and (persistent or auto or step)" 2>&1 | tail -4
} 2>&1 | tee gpurun_out/r2_k.txt
