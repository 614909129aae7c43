#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/s2j_pytest.log 2>&1; tail -8 gpurun_out/s2j_pytest.log
