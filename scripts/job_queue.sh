#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "queue or large_grid or edge_cases" 2>&1 | tail -4
for k in queue fused; do
  timeout 300 python bench.py --steps 8 --warmup 3 --kernel $k --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$k exact', d['ms_per_step'], d['roofline']['frac'])"
done
timeout 300 python bench.py --steps 8 --warmup 3 --kernel queue --mode fast --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('queue fast', d['ms_per_step'], d['roofline']['frac'])"
