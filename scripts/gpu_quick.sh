#!/bin/bash
# Run on the GPU box (gpurun): GPU parity tests + one bench line (no ncu). Usage: bash scripts/gpu_quick.sh [tag] [pytest-args]
tag=${1:-q}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu ${2:-} 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_$tag.log
python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -3 gpurun_out/pytest_gpu_$tag.log
tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'lat',d['latency'])
for k,v in d['kernels'].items(): print(f"{k:18s} {v['ms_per_step']*1000:7.1f} us {v['share']*100:5.1f}%")
PY
