#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_evaluation.py tests/test_callers.py -x -q -m gpu 2>&1 | tail -5 )
{ python scripts/e2e_probe.py 4096 1; python scripts/e2e_probe.py 4096 0; CC_B200_H2D_SPLIT=2 python scripts/e2e_probe.py 4096 1; CC_B200_BLOCKING_WAIT=1 python scripts/e2e_probe.py 4096 1; python scripts/e2e_probe.py 2048 1; python scripts/e2e_probe.py 6144 1; } 2>&1 | grep "rep [12]" | tee gpurun_out/e2e_probe.txt
python scripts/trace_push.py 4096 1 2>&1 | tail -45 > gpurun_out/tl4096_f9.txt; grep -A40 "push 11" gpurun_out/tl4096_f9.txt | grep -v "^  \." 
