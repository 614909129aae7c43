#!/bin/bash
# round 2, what is left of the GPU budget (1 GPU): the one-rank stress symmetrisation test on the last tree
mkdir -p gpurun_out
{ timeout 25 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tripole_stresses_resident_and_symmetrised" 2>&1 | tail -2; } 2>&1 | tee gpurun_out/r2_z8.txt
