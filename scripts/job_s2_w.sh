#!/bin/bash
# session 2, job W (last): full GPU suite, CD-grid timing
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/s2w_pytest.log 2>&1; tail -4 gpurun_out/s2w_pytest.log | cut -c1-200
timeout 120 python scripts/cdgrid_time.py 600 2>&1 | tail -1
