#!/bin/bash
# round 2, multi-GPU job (N = 2, 4 or 8 GPUs of one box; pass N as $1): parity of the in-kernel NVLink halo with the tile table in
# constant memory (written at the end of round 1, never run), then the weak-scaling bench line with it off and on.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- bash scripts/job_r2_scaling.sh 2
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
{
echo "== parity on $N GPUs, tile table in constant memory"
EVP_B200_P2P_CONST_TILES=1 run 29740 tests/mgpu_check.py gx1 40 48 120 fused 2>&1 | grep -E "MGPU|differs|rror" | cut -c1-200 | head -4
EVP_B200_P2P_CONST_TILES=1 run 29741 tests/mgpu_check.py gx3 10 10 30 fused - elim 2>&1 | grep -E "MGPU|differs|rror" | cut -c1-200 | head -4
echo "== parity on $N GPUs, two-lanes-per-cell kernel with the in-kernel halo (variant 40)"
EVP_B200_FUSED_VARIANT=40 run 29744 tests/mgpu_check.py gx1 40 48 120 fused 2>&1 | grep -E "MGPU|differs|rror" | cut -c1-200 | head -4
echo "== weak scaling, gx1-sized sub-domain per GPU"
timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2_scale_n1.json
run 29742 bench.py --gpus $N --steps 8 --warmup 3 2>gpurun_out/r2_scale_n$N.err | tail -1 > gpurun_out/r2_scale_n$N.json
EVP_B200_P2P_CONST_TILES=1 run 29743 bench.py --gpus $N --steps 8 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_scale_n${N}_ctiles.json
EVP_B200_FUSED_VARIANT=54 run 29746 bench.py --gpus $N --steps 8 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_scale_n${N}_v54.json
EVP_B200_FUSED_VARIANT=40 run 29745 bench.py --gpus $N --steps 8 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_scale_n${N}_v40.json
for f in gpurun_out/r2_scale_n${N}_v54.json gpurun_out/r2_scale_n${N}_v40.json gpurun_out/r2_scale_n1.json gpurun_out/r2_scale_n$N.json gpurun_out/r2_scale_n${N}_ctiles.json; do
  python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value', '%.3e'%d['value'], 'e2e', '%.3e'%d['e2e']['value'], d['config']['layout'][-100:])
"
done
tail -3 gpurun_out/r2_scale_n$N.err
} 2>&1 | tee gpurun_out/r2_scaling_n$N.txt
