#!/bin/bash
# compute-sanitizer passes over small lattice / band shapes.  Run on the GPU box:
#   gpurun --timeout 1500 -- bash scripts/gpu_sanitize.sh
# Logs land in gpurun_out/sanitize_*.log; summarise into profiles/ with scripts/summarise_sanitizer.py.
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL_DP='(chain or scan) and (shape0 or shape2) and ragged'
SEL_PATH='test_pipeline_vs_oracle and 5 or test_golden_simple_and_smoothed and c1 or test_band_recursion_with_large_delay_penalty and 400'
for tool in initcheck racecheck memcheck synccheck; do
  timeout 600 $SAN --tool $tool --error-exitcode 0 --log-file gpurun_out/sanitize_${tool}_dp.log \
      python -m pytest tests/test_gpu_dp.py -q -x -k "$SEL_DP" > gpurun_out/sanitize_${tool}_dp.out 2>&1
  echo "$tool dp rc=$?" >> gpurun_out/sanitize_rc.txt
  timeout 600 $SAN --tool $tool --error-exitcode 0 --log-file gpurun_out/sanitize_${tool}_path.log \
      python -m pytest tests/test_gpu_path.py -q -x -k "$SEL_PATH" > gpurun_out/sanitize_${tool}_path.out 2>&1
  echo "$tool path rc=$?" >> gpurun_out/sanitize_rc.txt
done
