#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_facade.py tests/test_callers.py -x -q -m gpu 2>&1 | tail -3 )
python bench.py --no-cpu-baseline > gpurun_out/bench_fac.json 2> gpurun_out/bench_fac.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_fac.json'))
for k,v in d['facade'].items(): print(k, round(v['columns_per_s']/1e3), 'k col/s', {a:round(b,1) for a,b in v.items() if a.endswith('p50')})
PY
