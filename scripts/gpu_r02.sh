#!/bin/bash
# GPU box, round 2: GPU test-suite, bench line (+ reference arm), device timelines (4096- and 64-firing pushes, visited-fix
# grid A/B), ncu launch list and one `--set full` capture of every kernel of a warm push.
# Usage: bash scripts/gpu_r02.sh [tag] [skip-tests]
tag=${1:-r02}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu_$tag.txt
if [ -z "$2" ]; then
  ( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_$tag.log
  tail -3 gpurun_out/pytest_gpu_$tag.log
fi
CC_BENCH_SLOT_TIMES=1 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 5 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
python scripts/trace_push.py 4096 > gpurun_out/tl4096_$tag.txt 2>&1
python scripts/trace_push.py 64 0 > gpurun_out/tl64_$tag.txt 2>&1
python scripts/trace_push.py 1024 0 wall > gpurun_out/tlwall_$tag.txt 2>&1
# the other BASELINE configurations (config 3: 128 rings; config 1 stand-in: 64 x 2200; moving sensor)
python bench.py --spec vls128 --no-cpu-baseline > gpurun_out/bench_vls128_$tag.json 2>> gpurun_out/bench_$tag.err
python bench.py --spec kitti64 --no-cpu-baseline --quick-e2e > gpurun_out/bench_kitti64_$tag.json 2>> gpurun_out/bench_$tag.err
python bench.py --moving --no-cpu-baseline > gpurun_out/bench_moving_$tag.json 2>> gpurun_out/bench_$tag.err
python scripts/e2e_timeline.py 4096 > gpurun_out/e2e_tl_$tag.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 4 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k_' -s 120 -c 16 \
    -o gpurun_out/prof_$tag -f python bench.py --steps 4 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'])
print('latency_mode', {k:v for k,v in d['latency_mode'].items() if k not in ('histogram_us','call')})
print('exact', d['exact_path'])
print('cpu', d['cpu_baseline']['value'])
for k,v in d['kernels'].items(): print(f"{k:18s} {v['ms_per_step']*1000:7.1f} us {v['share']*100:5.1f}%")
PY
for c in vls128 kitti64 moving; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${c}_$tag.json'))
    print('$c value',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),'lat', d.get('latency_mode') and round(d['latency_mode']['per_push_us_p50'],1))
except Exception as e: print('$c', e)
PY
done
tail -8 gpurun_out/e2e_tl_$tag.txt
