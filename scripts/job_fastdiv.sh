#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_cgrid.py -m gpu -q -x -k "not multi_gpu and not large_grid" 2>&1 | tail -4
for k in fused split persistent; do
  python bench.py --steps 8 --warmup 3 --kernel $k --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$k exact', d['ms_per_step'], d['roofline']['frac'])"
done
python bench.py --steps 8 --warmup 3 --kernel fused --mode fast --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fused fast', d['ms_per_step'], d['roofline']['frac'])"
