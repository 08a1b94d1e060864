#!/bin/bash
# GPU box: A/B of the result path of the end-to-end leg (same box, alternating): copy engine vs export kernel (grid sizes)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for rep in 1 2; do
for v in "CC_B200_RESULT_COPIES=1" "CC_B200_EXPORT_GRID=148" "CC_B200_EXPORT_GRID=16" "CC_B200_EXPORT_GRID=4"; do
  env $v CC_BENCH_SLOT_TIMES=1 python bench.py --no-cpu-baseline --quick-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
  echo "== $v"
  grep "e2e push [1-4]" gpurun_out/ab.err | cut -c1-120
  python - <<PY
import json
d=json.load(open('gpurun_out/ab.json'))
m=d['e2e']['wait_return_ms']
print('e2e',round(d['e2e']['value']/1e6,2),'first',m[0],'period',round((m[-1]-m[0])/(len(m)-1),4), 'value', round(d['value']/1e6,2))
PY
done
done
