#!/bin/bash
# Launch list + one `ncu --set full` capture of a c2 step (B200_PROFILING.md recipe).  usage: gpu_profile_step.sh TAG [workload]
set -u
TAG=${1:-rXX}; WL=${2:-c2}
mkdir -p gpurun_out
ARGS="--workload $WL --steps 3 --warmup 3 --no-graph --no-cpu --no-ref-gpu"
python bench.py $ARGS > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
python scripts/summarise_launches.py gpurun_out/${TAG}_launches.csv 60 > gpurun_out/${TAG}_launches_summary.txt
# full capture of one step's kernels (skip the warm-up launches)
ncu --set full --clock-control none --import-source on --launch-skip 60 -c 14 -f -o gpurun_out/${TAG}_step \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu2.log 2>&1
python scripts/ncu_summary.py gpurun_out/${TAG}_step.ncu-rep gpurun_out/${TAG}_ncu_full_step_summary.csv
cat gpurun_out/${TAG}_launches_summary.txt
