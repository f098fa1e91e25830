#!/bin/bash
# compute-sanitizer passes over every kernel family (run on the GPU box through gpurun):  scripts/sanitize_gpu.sh <tag>
tag=$1
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_workload.py \
      > gpurun_out/${tag}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?"; tail -n 4 gpurun_out/${tag}_sanitizer_${tool}.log
done
