#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: per kernel, instruction mix and the hottest SASS lines.
    python scripts/hot_sass.py gpurun_out/score_d10_source.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        body = []
        j = i + 2
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) >= len(hdr):
                body.append(rows[j])
            j += 1
        tot = sum(int(r[iS]) for r in body) or 1
        totI = sum(int(r[iI]) for r in body) or 1
        print("## %s\n   SASS lines %d, samples %d, warp-instructions %d" % (name[:100], len(body), tot, totI))
        ops = collections.Counter()
        smp = collections.Counter()
        for r in body:
            tok = r[1].split()
            op = tok[1] if tok and tok[0].startswith("@") else (tok[0] if tok else "?")
            ops[op.split(".")[0]] += int(r[iI])
            smp[op.split(".")[0]] += int(r[iS])
        print("   executed by opcode: " + ", ".join("%s %.1f%%" % (k, 100 * v / totI) for k, v in ops.most_common(22)))
        print("   samples by opcode:  " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in smp.most_common(14)))
        for r in sorted(body, key=lambda r: -int(r[iS]))[:top_n]:
            print("   %6s %5.1f%%  exec %9s  thr/inst %5.1f  %s" % (r[iS], 100 * int(r[iS]) / tot, r[iI],
                                                                  int(r[iT]) / max(1, int(r[iI])), r[1].strip()[:100]))
        i = j
    else:
        i += 1
