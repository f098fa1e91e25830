#!/bin/bash
# Run on the GPU box (gpurun), one part per call (gpurun copies back at most 64 MiB and a --set full capture is ~35 MB):
#   scripts/profile_gpu.sh <tag> launches|full|pscan [bench args...]
set -u
tag=$1; part=$2; shift 2
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency --prewarm-s 0"
case $part in
launches)  # launch list of the bench step
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
      $B "$@" > gpurun_out/${tag}_launches_bench.log 2>&1 ;;
full)      # one tick = linearize3, lq3 (structured) / lq_seq (dense), trial3 (alpha = 1), trial_reduce, wb_trial (2nd pass, warp form), trial3 (2nd pass),
           # trial_reduce, select_update, shift_extract: capture one full tick after a few warm ticks
  ncu --set full --clock-control none --import-source on -k regex:"linearize|lq3|lq_seq|lq_scan|trial|select_update" --launch-skip 24 -c 8 \
      -f -o gpurun_out/${tag}_full $B "$@" > gpurun_out/${tag}_full_bench.log 2>&1 ;;
pscan)     # parallel-in-time LQ (Talos single instance): elements, the first (full) and last (terminal-only) combine level, gains, rollout
  ncu --set full --clock-control none --import-source on -k regex:"pscan" --launch-skip 32 -c 8 \
      -f -o gpurun_out/${tag}_pscan_talos $B --workload talos_wb --envs 1 "$@" > gpurun_out/${tag}_pscan_talos_bench.log 2>&1 ;;
esac
ls -la gpurun_out | tail -5
