#!/bin/bash
tag=${1:-e2e}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python scripts/h2d_probe2.py 2>&1 | tee gpurun_out/h2d2_$tag.txt
python scripts/e2e_timeline.py 4096 2>&1 | tail -6 | tee gpurun_out/e2e_tl_$tag.txt
CC_B200_BLOCKING_WAIT=1 python scripts/e2e_timeline.py 4096 2>&1 | tail -6 | tee -a gpurun_out/e2e_tl_$tag.txt
python bench.py --quick --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'])
for k,v in d['kernels'].items(): print(f"{k:18s} {v['ms_per_step']*1000:7.1f} us {v['share']*100:5.1f}%")
PY
