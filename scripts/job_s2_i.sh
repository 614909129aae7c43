#!/bin/bash
# session 2, job I: resident stresses + 2-stream staging: full suite, bench lines
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s2i_pytest.log 2>&1; tail -15 gpurun_out/s2i_pytest.log
for w in gx1 gx3 tx1; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --workload $w 2>/dev/null | tail -1 > gpurun_out/s2i_bench_$w.json
  python -c "import json; d=json.load(open('gpurun_out/s2i_bench_$w.json')); print('$w', d['ms_per_step'], round(d['roofline']['frac'],4), 'e2e', d['e2e']['ms_per_step'], 'full', d['e2e_full_copy']['ms_per_step'])"
done
