#!/bin/bash
# GPU box: what the driver runs at round end -- the GPU test-suite, smoke(), the bench line and the reference arm
tag=${1:-final}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/pytest_gpu_$tag.log
tail -2 gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
r=json.loads([l for l in open('gpurun_out/bench_ref_$tag.json').read().splitlines() if l.startswith('{')][-1])
lm=d['latency_mode']
print('value',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),'ref',round(r['value']/1e6,4),'e2e ratio',round(d['e2e']['value']/r['value'],1),'value ratio',round(d['value']/r['value'],1))
print('latency p50/p99', round(lm['per_push_us_p50'],1), round(lm['per_push_us_p99'],1), 'dev', round(lm['per_push_device_us_p50'],1), 'same config', d['config']==r['config'])
PY
