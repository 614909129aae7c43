#!/bin/bash
# round 2, GPU job L (8 GPUs of one box): configs[4] (3600x2400, strong layout 4x2, 900x1200 per GPU), weak scaling at gx1 per GPU, and
# the per-GPU-size sweep of SURVEY 8(d); every line carries the bitwise parity pre-check against the CPU oracle.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- bash scripts/job_r2_l.sh
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
{
nvidia-smi -L | head -8
run 8 29811 bench.py --gpus 8 --workload p1deg --steps 5 --warmup 3 > gpurun_out/r2l_p1deg_n8.json 2> gpurun_out/r2l_p1deg_n8.err
run 8 29812 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2l_gx1_n8.json 2> gpurun_out/r2l_gx1_n8.err
run 8 29813 bench.py --gpus 8 --steps 10 --warmup 3 --sub 160x192 > gpurun_out/r2l_sub160_n8.json 2> gpurun_out/r2l_sub160_n8.err
run 8 29814 bench.py --gpus 8 --steps 6 --warmup 3 --sub 640x768 > gpurun_out/r2l_sub640_n8.json 2> gpurun_out/r2l_sub640_n8.err
run 8 29815 bench.py --gpus 8 --steps 10 --warmup 3 --kernel fused --no-parity > gpurun_out/r2l_gx1_n8_fused.json 2> gpurun_out/r2l_gx1_n8_fused.err
for f in p1deg_n8 gx1_n8 sub160_n8 sub640_n8 gx1_n8_fused; do python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2l_$f.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('$f', 'N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value %.3e'%d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity') and d['parity'].get('ok'), d['gpu_launches'], d['clocks'], '|', d['config']['workload'][:60], '|', d['config']['layout'][-150:])
except Exception as e:
    print('$f FAILED', e); print(open('gpurun_out/r2l_$f.err').read()[-2500:])
P
done
} 2>&1 | tee gpurun_out/r2_l.txt
