#!/bin/bash
# Run on the GPU box (gpurun): launch list + one ncu --set full capture of each hot kernel of the bench step.
#   scripts/profile_gpu.sh <tag> [bench args...]
set -u
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency --prewarm-s 0 "$@" > gpurun_out/${tag}_launches_bench.log 2>&1
# one tick = linearize3, lq_seq, trial3 (alpha = 1), trial_reduce, wb_trial (2nd pass, warp form), trial3 (2nd pass),
# trial_reduce, select_update, shift_extract: capture one full tick after a few warm ticks
ncu --set full --clock-control none --import-source on -k regex:"linearize|lq_seq|lq_scan|trial|select_update" --launch-skip 24 -c 8 \
    -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency --prewarm-s 0 "$@" > gpurun_out/${tag}_full_bench.log 2>&1
ls -la gpurun_out
# parallel-in-time LQ (Talos single instance): one full tick's pscan kernels
ncu --set full --clock-control none --import-source on -k regex:"pscan" --launch-skip 32 -c 8 \
    -f -o gpurun_out/${tag}_pscan_talos python bench.py --workload talos_wb --envs 1 --steps 2 --warmup 3 --no-cpu-baseline --no-latency --prewarm-s 0 > gpurun_out/${tag}_pscan_talos_bench.log 2>&1
ls -la gpurun_out | tail -5
