#!/bin/bash
# Run on the GPU box: the round's evidence set (tests, smoke, bench lines for the five BASELINE configs) -> gpurun_out/<tag>_*
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_gpu_tests.log 2>&1; tail -2 gpurun_out/${tag}_gpu_tests.log
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench_go2x1024.json 2> gpurun_out/${tag}_bench.err
python bench.py --workload h1_wb --envs 4096 --no-cpu-baseline --no-latency --no-strong > gpurun_out/${tag}_bench_h1x4096.json 2>> gpurun_out/${tag}_bench.err
python bench.py --workload srbd --envs 1 --steps 200 --warmup 10 --no-cpu-baseline --no-latency --no-strong > gpurun_out/${tag}_bench_srbd_e1.json 2>> gpurun_out/${tag}_bench.err
python bench.py --workload aliengo_wb --envs 1 --steps 200 --warmup 10 --no-cpu-baseline --no-latency --no-strong > gpurun_out/${tag}_bench_aliengo_wb_e1.json 2>> gpurun_out/${tag}_bench.err
python bench.py --workload talos_wb --envs 1 --steps 200 --warmup 10 --no-cpu-baseline --no-latency --no-strong > gpurun_out/${tag}_bench_talos_wb_e1.json 2>> gpurun_out/${tag}_bench.err
for f in gpurun_out/${tag}_bench_*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items()}, 'e2e', round(d['e2e']['value'],1), (d.get('latency') or {}).get('p50_ms'))"; done
