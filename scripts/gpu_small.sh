#!/bin/bash
# GPU box: the full bench line + device timelines (for profiles/)
tag=${1:-s}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CC_BENCH_SLOT_TIMES=1 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 5 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
python scripts/trace_push.py 4096 > gpurun_out/tl4096_$tag.txt 2>&1
python scripts/trace_push.py 64 0 > gpurun_out/tl64_$tag.txt 2>&1
python scripts/trace_push.py 1024 0 wall > gpurun_out/tlwall_$tag.txt 2>&1
grep "^push" gpurun_out/tlwall_$tag.txt
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
lm=d['latency_mode']
print('value',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),'cpu', round(d['cpu_baseline']['value']/1e6,4), 'lat', round(lm['per_push_us_p50'],1), round(lm['per_push_us_p99'],1))
print('exact', d['exact_path'])
PY
