#!/bin/bash
# GPU box with 8 GPUs: the bench line at N=8 (eight independent streams, BASELINE config 5)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nproc; lscpu | grep -i "numa\|socket\|model name" | head -8
nvidia-smi topo -m 2>/dev/null | head -14
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -3 gpurun_out/bench_n8.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n8.json").read().splitlines() if l.startswith("{")][-1])
print("value", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), [round(x/1e6,2) for x in d["e2e"].get("per_rank_columns_per_s") or []], d["run"].get("cpu_binding"))
print("python loop e2e", round(d["e2e"]["python_loop"]["value"]/1e6,2))
print([(x["rank"], round(x["per_push_us_p50"],1), round(x["per_push_us_p99"],1)) for x in d["per_gpu_latency"]])
PY
