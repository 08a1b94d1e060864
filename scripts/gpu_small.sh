#!/bin/bash
# GPU box: bench lines without the CPU baseline (default config + VLS-128)
tag=${1:-s}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CC_BENCH_SLOT_TIMES=1 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --spec vls128 --no-cpu-baseline > gpurun_out/bench_vls128_$tag.json 2>> gpurun_out/bench_$tag.err
tail -4 gpurun_out/bench_$tag.err
for f in bench_$tag bench_vls128_$tag; do python - <<PY
import json
d=json.load(open('gpurun_out/$f.json'))
lm=d['latency_mode']
print('$f value',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),'py',round(d['e2e']['python_loop']['value']/1e6,2),'ms/step',round(d['ms_per_step'],4))
print('  latency p50/p99', round(lm['per_push_us_p50'],1), round(lm['per_push_us_p99'],1), 'dev', round(lm['per_push_device_us_p50'],1), 'python', lm.get('python_loop'))
print('  sweep', {k:round(v['columns_per_s']/1e6,2) for k,v in (d.get('batch_sweep') or {}).items()})
PY
done
