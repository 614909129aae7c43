#!/bin/bash
# round 2, the last GPU seconds (2 GPUs): the stress symmetrisation between two ranks once more, now that the exchange takes its
# segment list from the host-side plan (evp_b200_stress_fold_plan) that the CPU suite executes over gloo
mkdir -p gpurun_out
{
export EVP_B200_P2P_TIMEOUT_S=2
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu and tiny-tripole-resident-stresses and nvlink" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2_z7.txt
