#!/bin/bash
# GPU box with 2 GPUs: the bench line at N=2 (two independent streams) and BASELINE config 4 (left / right OS-32)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --spec os32_pair --steps 20 --warmup 5 > gpurun_out/bench_os32pair.json 2>> gpurun_out/bench_n2.err
python - <<PY
import json
for f in ("bench_n2","bench_os32pair"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().splitlines() if l.startswith("{")][-1])
        print(f, "value", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), d["e2e"].get("per_rank_columns_per_s"), d["run"].get("cpu_binding"), [ (x["rank"], round(x["per_push_us_p50"],1)) for x in d["per_gpu_latency"]])
    except Exception as e: print(f, "ERR", e)
PY
