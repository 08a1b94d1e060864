#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_evaluation.py tests/test_callers.py -x -q -m gpu 2>&1 | tail -30 ) > gpurun_out/pytest_misc.log; tail -12 gpurun_out/pytest_misc.log
for g in 32 592; do echo "VFIX_GRID=$g"; CC_B200_VFIX_GRID=$g python scripts/trace_push.py 4096 1 2>&1 | grep -A45 "push 11" | grep "k_scan_lite\|k_scan_check\|k_fin_label\|k_visited_fix\|device_ms"; done
