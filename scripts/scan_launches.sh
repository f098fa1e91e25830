#!/bin/bash
# per-kernel times of one tick with the parallel-in-time LQ (ncu launch list):  scripts/scan_launches.sh <tag> <workload> [bench args]
tag=$1; wl=$2; shift 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_${wl}_scan_launches.csv \
  python bench.py --workload $wl --envs 1 --scan --steps 2 --warmup 3 --no-cpu-baseline --no-latency --prewarm-s 0 "$@" > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${tag}_${wl}_scan_launches.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    if "mpxb" not in r[4]: continue
    nm = r[4].split("(")[0][:60]; agg.setdefault(nm, []).append(float(r[-1].replace(",", "")))
for k, v in agg.items(): print("%-62s n=%3d mean %9.1f us" % (k, len(v), sum(v) / len(v) / 1e3))
PY
