#!/bin/bash
# GPU box: the full bench line
tag=${1:-s}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CC_BENCH_SLOT_TIMES=1 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -4 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
lm=d['latency_mode']
print('value',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),'py',round(d['e2e']['python_loop']['value']/1e6,2),'ms/step',round(d['ms_per_step'],4), 'cpu', round(d['cpu_baseline']['value']/1e6,4))
print('latency p50/p99', round(lm['per_push_us_p50'],1), round(lm['per_push_us_p99'],1), 'dev', round(lm['per_push_device_us_p50'],1))
print(json.dumps(d['rows_around_the_path'], indent=1))
PY
