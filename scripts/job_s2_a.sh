#!/bin/bash
# session 2, job A: full GPU suite, bench lines for fused / queue / C grid, smoke
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/s2a_pytest.log 2>&1
tail -25 gpurun_out/s2a_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
for k in fused queue persistent; do
  timeout 300 python bench.py --steps 8 --warmup 3 --kernel $k --no-cpu 2>/dev/null | tail -1 > gpurun_out/s2a_bench_$k.json
  python -c "import sys,json; d=json.load(open('gpurun_out/s2a_bench_$k.json')); print('$k exact', d['ms_per_step'], d['roofline']['frac'], d['e2e'])"
done
timeout 300 python bench.py --steps 8 --warmup 3 --kernel queue --mode fast --no-cpu 2>/dev/null | tail -1 > gpurun_out/s2a_bench_queue_fast.json
timeout 300 python bench.py --grid C --steps 4 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/s2a_bench_cgrid.json
cat gpurun_out/s2a_bench_cgrid.json
timeout 300 python bench.py --steps 3 --warmup 3 --workload p1deg --no-cpu 2>/dev/null | tail -1 > gpurun_out/s2a_bench_p1deg.json
cat gpurun_out/s2a_bench_p1deg.json
timeout 300 python bench.py --steps 3 --warmup 3 --workload p1deg --kernel queue --no-cpu 2>/dev/null | tail -1 > gpurun_out/s2a_bench_p1deg_queue.json
cat gpurun_out/s2a_bench_p1deg_queue.json
