#!/usr/bin/env python
"""Summarises what scripts/gpu_check.sh brought back in gpurun_out/ into profiles/<tag>_*.md|csv (tracked):
the ncu launch list (per-kernel device time and share of a step) and the `--set full` capture of the top kernels.
Usage: python scripts/summarize_ncu.py TAG"""
import csv
import collections
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(REPO, "profiles")
os.makedirs(out_dir, exist_ok=True)
g = os.path.join(REPO, "gpurun_out")
lines = [f"# ncu summary {tag}", ""]

launch_csv = os.path.join(g, f"launches_{tag}.csv")
if os.path.exists(launch_csv):
    rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 14 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        per.setdefault(name, []).append(float(r[14]))
    ours = {k: v for k, v in per.items() if k.startswith("k_")}
    tot = sum(sum(v) for v in ours.values())
    lines += [f"## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, `bench.py --steps 2 --warmup 3`)",
              "", f"{len(rows)} launches captured, {sum(len(v) for v in ours.values())} of them kernels of this library "
              "(the rest: torch fill/copy kernels of the bench harness). Times are cold-cache and serialised: compare shares.",
              "", "| kernel | launches | mean us | share of library time |", "|---|---|---|---|"]
    for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
    lines.append("")
    with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
        f.write("kernel,launches,mean_ns,share\n")
        for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k},{len(v)},{sum(v) / len(v):.0f},{sum(v) / tot:.4f}\n")

rep = os.path.join(g, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
    lines += ["## `ncu --set full --clock-control none --import-source on` of the top kernels (one launch each shown)", ""]
    seen = set()
    table = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        table[name] = {w: (r[idx[w]], units[idx[w]]) for w in want if w in idx}
    names = list(table)
    lines.append("| metric | " + " | ".join(names) + " |")
    lines.append("|---|" + "---|" * len(names))
    for w in want:
        if any(w in table[n] for n in names):
            lines.append(f"| {w} | " + " | ".join(f"{table[n][w][0]} {table[n][w][1]}" if w in table[n] else "" for n in names) + " |")
    lines.append("")
    json.dump(table, open(os.path.join(out_dir, f"{tag}_full.json"), "w"), indent=1)

    # dram__bytes_read.sum + dram__bytes_write.sum per launch: what bench.py reports as roofline.traffic
    def mbytes(v):
        val, unit = v
        val = float(val)
        return int(val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1))

    traffic = {"source": f"profiles/{tag}_full.json (ncu --set full --clock-control none, one launch per kernel, bench.py --quick at 4096 "
                         "firings per push; cold-cache replay)", "kernels": {}}
    for n, t in table.items():
        if "dram__bytes_read.sum" in t and "dram__bytes_write.sum" in t:
            rd, wr = mbytes(t["dram__bytes_read.sum"]), mbytes(t["dram__bytes_write.sum"])
            traffic["kernels"][n] = {"dram_bytes": rd + wr, "read": rd, "write": wr}
    if "k_fin_cluster" in traffic["kernels"]:  # the whole-push finish pass is timed under the name of its single-CTA variant
        traffic["kernels"]["k_fin_all"] = traffic["kernels"]["k_fin_cluster"]
    json.dump(traffic, open(os.path.join(out_dir, "ncu_traffic.json"), "w"), indent=1)

bench = os.path.join(g, f"bench_{tag}.json")
if os.path.exists(bench):
    txt = open(bench).read().strip().splitlines()
    if txt:
        lines += ["## bench line of the same gpurun call", "", "```json", txt[-1], "```", ""]
        open(os.path.join(out_dir, f"{tag}_bench.json"), "w").write(txt[-1] + "\n")
ref = os.path.join(g, f"bench_ref_{tag}.json")
if os.path.exists(ref):
    txt = [l for l in open(ref).read().strip().splitlines() if l.startswith("{")]
    if txt:
        lines += ["## reference arm (`bench.py --impl reference`) on the same box", "", "```json", txt[-1], "```", ""]
open(os.path.join(out_dir, f"{tag}_summary.md"), "w").write("\n".join(lines))
print("\n".join(lines[:40]))
