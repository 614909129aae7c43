#!/bin/bash
# session 2, job V: CD grid (a13) CUDA path: parity against the oracle and the reference-source vectors; C grid regression
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_cgrid.py -m gpu -q ) > gpurun_out/s2v_pytest.log 2>&1; tail -30 gpurun_out/s2v_pytest.log | cut -c1-220
