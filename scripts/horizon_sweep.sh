#!/bin/bash
# Talos whole-body single instance, horizon sweep with both LQ kernels (BASELINE configs[3]):  scripts/horizon_sweep.sh <tag>
tag=$1
mkdir -p gpurun_out
: > gpurun_out/${tag}_talos_horizon.jsonl
for N in 25 50 100 200; do
  for mode in "SEQ" "--scan --scan-schedule parallel" ${SWEEP_SERIAL:+"--scan --scan-schedule serial"}; do
    if [ "$mode" = "SEQ" ]; then export MPXB_LQ=sequential; mode=""; else unset MPXB_LQ; fi
    python bench.py --workload talos_wb --envs 1 --horizon $N $mode --steps 100 --warmup 10 --no-cpu-baseline --no-latency \
      --prewarm-s 0.5 >> gpurun_out/${tag}_talos_horizon.jsonl 2>> gpurun_out/${tag}_talos_horizon.err
  done
done
python - <<PY
import json
for l in open("gpurun_out/${tag}_talos_horizon.jsonl"):
    d = json.loads(l)
    k = d["roofline"]["kernel_ms"]
    print(d["config"]["horizon"], d["config"]["lq"], "tick %.3f ms" % d["ms_per_step"], "lq %.3f ms" % k["lq_solve"], "lin %.3f" % k["linearize"], "finite", d["nonfinite_envs"] == 0)
PY
