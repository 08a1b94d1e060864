#!/bin/bash
# Run on the GPU box (gpurun): GPU parity tests, a bench line, the ncu launch list and one full capture of the top kernels.
# Usage: bash scripts/gpu_check.sh [tag]
tag=${1:-r01}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu_$tag.txt
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 5 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_insert_scan|k_probe|k_ground' -s 12 -c 6 \
    -o gpurun_out/prof_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
tail -3 gpurun_out/pytest_gpu_$tag.log
head -c 1200 gpurun_out/bench_$tag.json
