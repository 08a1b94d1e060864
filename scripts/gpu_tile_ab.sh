#!/bin/bash
# GPU box: parity of the tiled probe (bulk-async staged window) + A/B against the list-driven probe (CC_B200_TUNE=4)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "small" 2>&1 | tail -4 ) > gpurun_out/pytest_tile.log
tail -3 gpurun_out/pytest_tile.log
for v in 0 4 0 4; do
  CC_B200_TUNE=$v python bench.py --no-cpu-baseline --quick-e2e > gpurun_out/tile_$v.json 2> gpurun_out/tile_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/tile_$v.json'))
k=d['kernels']; tl=d['kernels_device_timeline_us']
print('tune=$v value',round(d['value']/1e6,2),'ms/step',round(d['ms_per_step'],4),'dev p50',round(d['latency']['per_push_device_ms_p50'],4),
      '| timeline probe', tl.get('k_probe'), tl.get('k_probe_tile'), 'heavy', tl.get('k_probe_heavy'), '| e2e', round(d['e2e']['value']/1e6,2))
PY
done
CC_B200_TUNE=0 python scripts/trace_push.py 4096 > gpurun_out/tl4096_tile.txt 2>&1
grep "device_ms\|k_probe" gpurun_out/tl4096_tile.txt | tail -6
