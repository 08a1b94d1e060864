#!/bin/bash
# GPU box: bench line without the CPU baseline + end-to-end probes
tag=${1:-s}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CC_BENCH_SLOT_TIMES=1 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -8 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'])
print('marks', d['e2e'].get('wait_return_ms'))
print('latency_mode', {k:v for k,v in d['latency_mode'].items() if k not in ('histogram_us','call')})
print('exact', d['exact_path'])
PY
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "staged or pipelined or full_size" 2>&1 | tail -3 ) 
