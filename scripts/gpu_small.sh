#!/bin/bash
# GPU box: the full bench line
tag=${1:-s}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CC_BENCH_SLOT_TIMES=1 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -3 gpurun_out/bench_$tag.err | grep -v "e2e push"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
lm=d['latency_mode']
print('value',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),'cpu', round(d['cpu_baseline']['value']/1e6,4), 'lat', round(lm['per_push_us_p50'],1), round(lm['per_push_us_p99'],1))
print('exact', d['exact_path']['columns_per_s'], 'irregular', d['irregular_stream'])
PY
