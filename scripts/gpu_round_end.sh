set -u
TAG=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; echo tests rc=$?; tail -1 gpurun_out/${TAG}_gputests.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err; echo ref rc=$?
python bench.py --workload c4 --steps 30 --warmup 5 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err; echo c4 rc=$?
python bench.py --workload c5 --steps 40 --warmup 5 > gpurun_out/${TAG}_c5_1gpu.json 2> gpurun_out/${TAG}_c5_1gpu.err; echo c5 rc=$?
scripts/gpu_profile_step.sh ${TAG} c2 > /dev/null 2>&1
scripts/gpu_profile_step.sh ${TAG}_c4 c4 > /dev/null 2>&1
python -c "
import json
for f in ('bench','c4','c5_1gpu','reference_arm'):
    d=json.loads(open('gpurun_out/${TAG}_'+f+'.json').read().strip().splitlines()[-1]); print(f, d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'), d.get('stages_ms'))
"
cat gpurun_out/${TAG}_launches_summary.txt
