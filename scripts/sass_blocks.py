#!/usr/bin/env python
"""Hot basic blocks of a kernel from an `ncu --page source --csv` dump: contiguous SASS runs with one execution count,
with their share of the stall samples and of the executed warp instructions.

    python scripts/sass_blocks.py gpurun_out/tc_topk_d10_source.csv [top_n] [n_warps]
"""
import collections
import csv
import sys

csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
n_warps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
k = []
for r in rows[2:]:
    if r and r[0].startswith("Kernel Name"):
        break
    if len(r) >= 11 and r[0].startswith("0x"):
        k.append(r)
tot_s = sum(int(r[2]) for r in k)
tot_i = sum(int(r[5]) for r in k)
print(rows[0][1])
print(len(k), "SASS instructions; stall samples", tot_s, "; warp instructions executed", tot_i)
blocks, cur = [], None
for idx, r in enumerate(k):
    ex = int(r[5])
    if cur is None or ex != cur["ex"]:
        cur = {"start": idx, "ex": ex, "n": 0, "s": 0, "ops": []}
        blocks.append(cur)
    cur["n"] += 1
    cur["s"] += int(r[2])
    t = r[1].split()
    cur["ops"].append(t[1] if t[0].startswith("@") else t[0])
blocks.sort(key=lambda b: -b["ex"] * b["n"])
for b in blocks[:top_n]:
    c = " ".join("%s:%d" % kv for kv in collections.Counter(b["ops"]).most_common(7))
    print("sass %5d +%4d  exec/inst %9.1f  samples %5.1f%%  inst %5.1f%%  %s"
          % (b["start"], b["n"], b["ex"] / n_warps, 100.0 * b["s"] / tot_s, 100.0 * b["ex"] * b["n"] / tot_i, c))
