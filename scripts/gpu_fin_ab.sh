#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "small or staged or full_size or golden" 2>&1 | tail -3 )
python scripts/trace_push.py 4096 > gpurun_out/tl4096_fin.txt 2>&1
grep "device_ms\|k_fin_all\|fin_" gpurun_out/tl4096_fin.txt | tail -16
python bench.py --no-cpu-baseline --quick-e2e > gpurun_out/fin.json 2> gpurun_out/fin.err
python - <<PY
import json
d=json.load(open('gpurun_out/fin.json'))
print('value',round(d['value']/1e6,2),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']/1e6,2), d['kernels_device_timeline_us'])
PY
