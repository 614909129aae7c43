#!/bin/bash
# round 2, GPU job E (1 GPU): the rewritten persistent kernel -- parity tests, bench at gx1 beside the default kernel
mkdir -p gpurun_out
{
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "persistent" 2>&1 | tail -15
for kern in persistent auto; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-pageable --kernel $kern > gpurun_out/r2e_gx1_$kern.json 2> gpurun_out/r2e_gx1_$kern.err
  python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2e_gx1_$kern.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('$kern', 'ms/step', round(d['ms_per_step'],3), 'us/subcycle', round(d['roofline']['us_per_subcycle'],3), 'frac', round(d['roofline']['frac'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity') and d['parity'].get('ok'), '|', d['config']['layout'][-200:])
except Exception as e:
    print('$kern FAILED', e); print(open('gpurun_out/r2e_gx1_$kern.err').read()[-1500:])
P
done
timeout 300 python bench.py --workload gx3 --steps 10 --warmup 3 --no-cpu --no-pageable --kernel persistent | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gx3 persistent ms/step', d['ms_per_step'], d['parity'])"
timeout 300 python bench.py --workload gx3 --steps 10 --warmup 3 --no-cpu --no-pageable | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gx3 auto ms/step', d['ms_per_step'], d['parity'])"
} 2>&1 | tee gpurun_out/r2_e.txt
