#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_callers.py tests/test_facade.py -x -q -m gpu -k "small or cuda or full_size" 2>&1 | tail -3 )
python scripts/trace_push.py 4096 > gpurun_out/tl4096_vfix.txt 2>&1
grep "device_ms\|k_visited_fix\|k_fin_label" gpurun_out/tl4096_vfix.txt | tail -6
python bench.py --no-cpu-baseline --quick-e2e > gpurun_out/vfix.json 2> gpurun_out/vfix.err
python - <<PY
import json
d=json.load(open('gpurun_out/vfix.json'))
print('value',round(d['value']/1e6,2),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']/1e6,2))
PY
