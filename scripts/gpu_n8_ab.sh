#!/bin/bash
# 8 GPUs: what bounds the device-resident leg at N=8? result path (export kernel vs copy engine) x CPU binding
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
port=29530
for v in "A=1" "CC_B200_RESULT_COPIES=1" "CC_BENCH_NO_BIND=1" "CC_B200_RESULT_COPIES=1 CC_BENCH_NO_BIND=1"; do
  port=$((port+1))
  env $v python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 20 --warmup 5 --quick-e2e > gpurun_out/n8ab.json 2> gpurun_out/n8ab.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/n8ab.json").read().splitlines() if l.startswith("{")][-1])
print("$v", "| value", round(d["value"]/1e6,1), "e2e", round(d["e2e"]["value"]/1e6,1), [round(x/1e6,1) for x in d["e2e"].get("per_rank_columns_per_s") or []], "py", round(d["e2e"]["python_loop"]["value"]/1e6,1))
PY
done
