#!/bin/bash
# usage (on the GPU box): scripts/gpu_multi.sh N TAG  -- c2 (weak scaling) and c5 (one ragged batch, strong scaling) bench lines on N GPUs
N=$1; TAG=$2; mkdir -p gpurun_out
run() { if [ "$N" = 1 ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; fi; }
run --steps 200 --warmup 5 --no-training > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "c2 rc=$?"
run --workload c5 --steps 40 --warmup 5 > gpurun_out/${TAG}_c5_${N}gpu.json 2> gpurun_out/${TAG}_c5_${N}gpu.err; echo "c5 rc=$?"
tail -c 600 gpurun_out/${TAG}_bench_${N}gpu.json; tail -c 400 gpurun_out/${TAG}_c5_${N}gpu.json
