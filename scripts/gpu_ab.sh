#!/bin/bash
# A/B of CC_B200_TUNE values on the GPU box: bash scripts/gpu_ab.sh "0 1"
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for t in $1; do
  CC_B200_TUNE=$t timeout 120 python bench.py --quick --no-cpu-baseline > gpurun_out/bench_ab$t.json 2> gpurun_out/bench_ab$t.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab$t.json'))
print('tune=$t value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],4),'dev p50',round(d['latency']['per_push_device_ms_p50'],4))
print('   '+' '.join(f"{k[2:]}={v['ms_per_step']*1000:.1f}" for k,v in d['kernels'].items()))
PY
done
