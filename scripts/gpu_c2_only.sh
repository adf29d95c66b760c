#!/bin/bash
# usage (on the GPU box): scripts/gpu_c2_only.sh N TAG -- the c2 (weak scaling) line on N GPUs
N=$1; TAG=$2; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 200 --warmup 5 --no-training > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_${N}gpu.json').read().strip().splitlines()[-1]); print($N, d['ms_per_step'], d['value'], d['e2e']['value'])"
