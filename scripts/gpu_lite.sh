#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "small or staged or full_size or golden or pipelined" 2>&1 | tail -3 )
python scripts/trace_push.py 4096 > gpurun_out/tl4096_lite.txt 2>&1
grep "device_ms\|k_scan_lite\|lite_p1\|k_scan_check" gpurun_out/tl4096_lite.txt | tail -8
