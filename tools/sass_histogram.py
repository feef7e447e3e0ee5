#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libcnvetti_b200.so (cuobjdump -sass): the evidence for which
hardware paths each kernel uses (UTCHMMA / UTMALDG / LDTM = tcgen05 / TMA / TMEM, RED / ATOMS = atomics,
VIMNMX3 / VIADDMNMX = DPX, LDG.E.128 ...).  Writes profiles/r2_sass_opcodes.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "cnvetti_b200", "libcnvetti_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "")
        name = re.sub(r"^void\s+", "", name)
        base = re.split(r"[<(]", name, 1)[0]
        targs = re.match(r"[^<(]*<([^(]*)>\(", name)
        kern = base + ("<%s>" % targs.group(1) if targs and not base.startswith("cub::") else "")
        hist.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Za-z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
interesting = re.compile(r"^(LDGSTS|F2F|UTC|UTMA|LDTM|STTM|UBLKCP|SYNCS|RED|ATOM|ATOMS|ATOMG|VIMNMX|VIADDMNMX|FMNMX|LDG\.E\.128|LDG\.E\.64|STG\.E\.128|LDS\.128|STS\.128|SHFL|MATCH|VOTE|REDUX|HMMA|IMMA|DADD|DFMA|DMUL|POPC|LOP3|PRMT|BAR|ELECT|CCTL|FENCE|MEMBAR|UCGABAR)")
with open(os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt"), "w") as f:
    f.write("# cuobjdump -sass cnvetti_b200/libcnvetti_b200.so (sm_100a), opcode counts per kernel; tools/sass_histogram.py\n")
    for k, c in hist.items():
        total = sum(c.values())
        keys = sorted((o for o in c if interesting.match(o)), key=lambda o: -c[o])
        f.write("\n%s   [%d instructions]\n" % (k, total))
        f.write("  " + "  ".join("%s:%d" % (o, c[o]) for o in keys) + "\n")
        top = ", ".join("%s:%d" % (o, n) for o, n in c.most_common(8))
        f.write("  top: " + top + "\n")
print("kernels:", len(hist))
