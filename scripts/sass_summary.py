#!/usr/bin/env python
"""SASS evidence for the tensor-core / TMA path: per kernel of liblec_b200.so, counts of the mnemonics that prove
tcgen05 (UTCHMMA, UTCBAR, LDTM, STTM, UTCATOMSWS), bulk copies (UBLKCP), mbarriers (SYNCS), packed fp32 (FFMA2 ...),
vector reductions (REDG) and the register / spill figures ptxas reported.

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "learning_embeddings_b200", "_lib", "liblec_b200.so")
KEYS = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "SYNCS", "ELECT", "FFMA2", "FMUL2", "FADD2", "MUFU",
        "REDG", "LDG", "STG", "DFMA", "SHFL", "ACQBULK", "NANOSLEEP")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        counts[kern][m.group(1)] += 1
        total[kern] += 1
print("SASS mnemonic counts per kernel of %s (sm_100a)\n" % os.path.relpath(LIB, ROOT))
for k, c in counts.items():
    short = re.sub(r"\(.*", "", k).replace("lec::", "")
    hits = ", ".join("%s %d" % (key, c[key]) for key in KEYS if c[key])
    print("%-58s %5d instr | %s" % (short[:58], total[k], hits))
