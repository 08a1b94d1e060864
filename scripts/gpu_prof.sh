#!/bin/bash
# Run on the GPU box (gpurun): bench line + ncu launch list + `--set full` capture of every kernel of one warm push.
# Usage: bash scripts/gpu_prof.sh [tag]
tag=${1:-p}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k_' -s 140 -c 14 \
    -o gpurun_out/prof_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'lat',d['latency'])
for k,v in d['kernels'].items(): print(f"{k:18s} {v['ms_per_step']*1000:7.1f} us {v['share']*100:5.1f}%")
PY
