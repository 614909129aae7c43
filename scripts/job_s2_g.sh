#!/bin/bash
# session 2, job G: cooperative single-launch C-grid kernel vs the five-kernel form; full suite with the new defaults
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s2g_pytest.log 2>&1; tail -3 gpurun_out/s2g_pytest.log
echo "cgrid 5-kernel parity: $(EVP_B200_CGRID_COOP=0 timeout 600 python -m pytest tests/test_cgrid.py -m gpu -x -q 2>&1 | tail -1)"
echo "--- coop"; timeout 300 python scripts/cgrid_time.py 600 2>&1 | tail -2
echo "--- five kernels"; EVP_B200_CGRID_COOP=0 timeout 300 python scripts/cgrid_time.py 600 2>&1 | tail -2
timeout 300 python bench.py --grid C --steps 4 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/s2g_bench_cgrid.json
python -c "import json; d=json.load(open('gpurun_out/s2g_bench_cgrid.json')); print('bench C:', d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['ms_per_step'])"
b() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['e2e']['ms_per_step'])"; }
echo "gx1 default: $(b)"
echo "p1deg default: $(b --workload p1deg --steps 3)"
