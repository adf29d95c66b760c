N=$1
for nb in 1 2 3 6; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload c5 --steps 40 --warmup 5 --buckets $nb --no-cpu > gpurun_out/r02ao_c5_${N}gpu_b$nb.json 2> gpurun_out/r02ao_c5_${N}gpu_b$nb.err
python -c "
import json; d=json.loads(open('gpurun_out/r02ao_c5_${N}gpu_b$nb.json').read().strip().splitlines()[-1]); print('buckets', $nb, d['ms_per_step'], d['value'])"
done
