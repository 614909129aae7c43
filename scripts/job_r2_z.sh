#!/bin/bash
# round 2, final GPU job (1 GPU): whole GPU suite, smoke, the default bench line (gx1) and the reference arm, the bench line of
# 3600x2400 on one GPU (AUTO -> TMA tile-streaming kernel), ncu launch list of that bench command, ncu --set full of the
# tile-streaming kernel in its final form
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | cut -c1-150
timeout 300 python bench.py > gpurun_out/r2z_bench_gx1.json 2> gpurun_out/r2z_bench_gx1.err; tail -2 gpurun_out/r2z_bench_gx1.err
timeout 300 python bench.py --workload p1deg --steps 3 --warmup 3 --no-cpu --no-pageable > gpurun_out/r2z_bench_p1deg.json 2> gpurun_out/r2z_bench_p1deg.err; tail -2 gpurun_out/r2z_bench_p1deg.err
timeout 300 python bench.py --workload p1deg --kernel stream --steps 3 --warmup 3 --no-cpu --no-pageable --no-parity > gpurun_out/r2z_bench_p1deg_fused.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench_p1deg.csv python bench.py --workload p1deg --steps 1 --warmup 1 --no-cpu --no-pageable --no-parity > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tstream_kernel -s 6 -c 1 -o gpurun_out/r2_tstream_final -f python scripts/prof_step.py p1deg tstream exact 4 3 > gpurun_out/r2z_ncu.log 2>&1; tail -2 gpurun_out/r2z_ncu.log
python - <<P
import json
for f in ('gx1','p1deg','p1deg_fused'):
    try:
        d=json.loads([l for l in open('gpurun_out/r2z_bench_%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.4e'%d['value'], 'ms', round(d['ms_per_step'],3), 'kernel', d['config']['kernel'], 'e2e', round(d['e2e'].get('ms_per_step',0),3), 'parity', (d.get('parity') or {}).get('ok'), 'clocks', d.get('clocks'), 'frac', (d.get('roofline') or {}).get('frac'))
    except Exception as e:
        print(f, 'FAILED', e)
P
} 2>&1 | tee gpurun_out/r2_z.txt
