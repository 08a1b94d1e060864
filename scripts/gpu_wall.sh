#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_edge_cases.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 )
python bench.py --no-cpu-baseline > gpurun_out/bench_wall.json 2> gpurun_out/bench_wall.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_wall.json'))
print('value',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2))
print('exact', d['exact_path'])
PY
