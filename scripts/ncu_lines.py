#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel from an ncu report (needs -lineinfo + --import-source on).
Usage: python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_NAME [TOP]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("CC_LIB", os.path.join(REPO, "continuous_clustering_b200", "libcc_b200.so"))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines())]
# several launches may be in the report: keep the first block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
    elif cur is not None and r and r[0].startswith("0x"):
        cur.append(r)
sass = blocks[0]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", os.path.join(td, cubin)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kernel in l)
lines, cur = [], None
for l in dis[start + 1:]:
    if l.startswith("\t.section") or (l.startswith(".text.") and kernel not in l):
        break
    m = re.search(r'//## File ".*?/([\w.]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
    elif re.search(r"/\*[0-9a-f]{4}\*/\s+\S", l):
        lines.append(cur)
assert len(lines) == len(sass), (len(lines), len(sass))
agg, inst = collections.Counter(), collections.Counter()
for r, src in zip(sass, lines):
    if r[2].isdigit():
        agg[src] += int(r[2])
        inst[src] += int(r[5]) if r[5].isdigit() else 0
tot = sum(agg.values())
print(f"{kernel}: {tot} samples, {sum(inst.values())} warp instructions")
cache = {}
for src, c in agg.most_common(top):
    text = ""
    if src:
        path = os.path.join(REPO, "continuous_clustering_b200", "csrc", src[0])
        if os.path.exists(path):
            cache.setdefault(path, open(path).read().splitlines())
            text = cache[path][src[1] - 1].strip()[:95]
    print(f"{(src[0] + ':' + str(src[1])) if src else '?':>22} {c:6d} {100 * c / tot:5.1f}% {inst[src]:9d}  {text}")
