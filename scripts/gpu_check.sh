#!/bin/bash
# Run on the GPU box (gpurun): GPU parity tests, the bench line and the reference arm, the device timeline of a push, the
# ncu launch list and one `--set full` capture of the top kernels.
# Usage: bash scripts/gpu_check.sh [tag]
tag=${1:-r01}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu_$tag.txt
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 5 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
python scripts/trace_push.py 4096 > gpurun_out/timeline_$tag.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 4 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_fin_all|k_ground|k_probe|k_prep|k_scan_lite|k_gap_scan' -s 30 -c 12 \
    -o gpurun_out/prof_$tag -f python bench.py --steps 4 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
tail -3 gpurun_out/pytest_gpu_$tag.log
head -c 600 gpurun_out/bench_$tag.json
