#!/bin/bash
# round 2, final single-GPU evidence: whole GPU suite, smoke, default bench line, ncu launch list of the bench command, ncu captures of
# the streaming form at 3600x2400 and of the C grid line
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | cut -c1-160
timeout 300 python bench.py > gpurun_out/r2f_bench_gx1.json 2> gpurun_out/r2f_bench_gx1.err; tail -2 gpurun_out/r2f_bench_gx1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_gx1.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-pageable --no-parity > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:fused_kernel -s 6 -c 1 -o gpurun_out/r2_stream_p1deg -f python scripts/prof_step.py p1deg stream exact 4 3 > gpurun_out/r2f_p1deg_ncu.log 2>&1; tail -2 gpurun_out/r2f_p1deg_ncu.log
timeout 300 python bench.py --grid C --steps 5 --warmup 3 > gpurun_out/r2f_bench_cgrid.json 2>/dev/null
timeout 300 python bench.py --workload tx1 --steps 10 --warmup 3 --no-cpu --no-pageable > gpurun_out/r2f_bench_tx1.json 2>/dev/null
python - <<P
import json
for f in ('gx1','cgrid','tx1','ref'):
    try:
        d=json.loads([l for l in open('gpurun_out/r2f_bench_%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.4e'%d['value'], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e'].get('ms_per_step',0),3), 'parity', (d.get('parity') or {}).get('ok'), 'clocks', d.get('clocks'), 'frac', (d.get('roofline') or {}).get('frac'))
    except Exception as e:
        print(f, 'FAILED', e)
P
} 2>&1 | tee gpurun_out/r2_final1.txt
