#!/bin/bash
# session 2, job U: full GPU suite (incl. the reference-source vectors) + smoke()
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/s2u_pytest.log 2>&1; tail -6 gpurun_out/s2u_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | cut -c1-120
