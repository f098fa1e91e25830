#!/bin/bash
# Run on the GPU box (gpurun): launch list + one ncu --set full capture of each hot kernel of the bench step.
#   scripts/profile_gpu.sh <tag> [bench args...]
set -u
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"linearize|lq_seq|lq_scan|trial" --launch-skip 12 -c 3 \
    -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_full_bench.log 2>&1
ls -la gpurun_out
