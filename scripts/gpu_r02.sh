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
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 5 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
for g in 32 296; do
  CC_B200_VFIX_GRID=$g python scripts/trace_push.py 4096 > gpurun_out/tl4096_${tag}_v$g.txt 2>&1
done
python scripts/trace_push.py 64 0 > gpurun_out/tl64_$tag.txt 2>&1
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
grep -h "device_ms\|k_visited_fix\|k_fin_all " gpurun_out/tl4096_${tag}_v*.txt | tail -12
tail -8 gpurun_out/e2e_tl_$tag.txt
