# A/B timing of library variants: scripts/quick_ab.sh "label:ENV=..;ENV=.." ...   (Go2 x1024 tick + per-kernel ms)
Q="--no-cpu-baseline --no-latency --no-strong --steps 20 --warmup 3"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["kernel_ms"].items()})'
for spec in "$@"; do
  label=${spec%%:*}; envs=${spec#*:}
  echo "== $label"
  ( IFS=';'; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done; unset IFS; python bench.py $Q $EXTRA 2>/dev/null | python -c "$pick" )
done
