#!/bin/bash
# usage (on the GPU box): scripts/gpu_c5_only.sh N TAG [extra bench args] -- the c5 line on N GPUs
N=$1; TAG=$2; shift 2; mkdir -p gpurun_out
if [ "$N" = 1 ]; then python bench.py --gpus 1 --workload c5 --steps 40 --warmup 5 "$@" > gpurun_out/${TAG}_c5_${N}gpu.json 2> gpurun_out/${TAG}_c5_${N}gpu.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c5 --steps 40 --warmup 5 "$@" > gpurun_out/${TAG}_c5_${N}gpu.json 2> gpurun_out/${TAG}_c5_${N}gpu.err; fi
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_c5_${N}gpu.json').read().strip().splitlines()[-1]); print($N, d['ms_per_step'], d['value'], d['e2e']['value'], len(d['config']['buckets_rank0']))"
