#!/bin/bash
# GPU box: parity tests, then latency probes and timelines of short pushes (fused kernel vs chain)
tag=${1:-lat}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_$tag.log
tail -3 gpurun_out/pytest_gpu_$tag.log
for b in 1 16 64 128 256 512; do python scripts/lat_probe.py $b; done 2>&1 | tee gpurun_out/lat_$tag.txt
for b in 64 128 256 512; do CC_B200_FUSED_MAX=0 python scripts/lat_probe.py $b; done 2>&1 | tee -a gpurun_out/lat_$tag.txt
python scripts/trace_push.py 64 0 > gpurun_out/tl64_$tag.txt 2>&1
python scripts/trace_push.py 256 0 > gpurun_out/tl256_$tag.txt 2>&1
grep -A60 "push 11" gpurun_out/tl64_$tag.txt | head -70
