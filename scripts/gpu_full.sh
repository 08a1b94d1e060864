#!/bin/bash
# GPU box: full GPU test-suite, the bench line (+ reference arm), e2e timeline
tag=${1:-full}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/pytest_gpu_$tag.log
tail -4 gpurun_out/pytest_gpu_$tag.log
python scripts/e2e_timeline.py 4096 > gpurun_out/e2e_tl_$tag.txt 2>&1
tail -12 gpurun_out/e2e_tl_$tag.txt
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'])
print('latency_mode', {k:v for k,v in d['latency_mode'].items() if k!='histogram_us'})
print('facade', json.dumps(d['facade'], indent=1))
print('exact', d['exact_path'])
print('cpu', d['cpu_baseline'])
print('roofline', {k:v for k,v in d['roofline'].items() if k!='by_kernel'})
for k,v in d['kernels'].items(): print(f"{k:18s} {v['ms_per_step']*1000:7.1f} us {v['share']*100:5.1f}%")
PY
