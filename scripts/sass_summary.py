#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libpawb200.so (cuobjdump -sass, no GPU needed) -> markdown table.

  python scripts/sass_summary.py [pawpyseed_b200/libpawb200.so] > profiles/r02_sass.md

Columns are the mnemonics that identify the data path of a kernel on sm_100a: DMMA (FP64 tensor core,
mma.sync.m8n8k4.f64), LDGSTS (cp.async), UTMALDG / UBLKCP (TMA tensor / bulk copies), SYNCS (mbarrier), DFMA/DADD/DMUL
(FP64 pipe), LDG/STG/LDS/STS, BAR, and the spill traffic LDL/STL."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "pawpyseed_b200/libpawb200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cols = ["DMMA", "LDGSTS", "UTMALDG", "UBLKCP", "UBLKPF", "CCTL", "SYNCS", "DFMA", "DADD", "DMUL", "LDG", "STG", "LDS", "STS", "BAR",
        "ATOM", "RED", "MEMBAR", "LDL", "STL"]
kern = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        kern[cur]["_total"] += 1
        for c in cols:
            if op == c or op.startswith(c + "."):
                kern[cur][c] += 1
names = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
print("# r02 SASS mnemonic counts per kernel (`cuobjdump -sass %s`, sm_100a)\n" % so)
print("Static instruction counts (not executed counts).  `DMMA` = FP64 tensor-core MMA (tcgen05 has no FP64 kind), "
      "`LDGSTS` = cp.async, `UTMALDG` = cp.async.bulk.tensor (TMA), `SYNCS` = mbarrier ops.\n")
print("| kernel | total | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
for (k, c), n in zip(kern.items(), names):
    n = re.sub(r"\(.*", "", n).replace("pawb200::", "").replace("void ", "")
    print("| `%s` | %d | " % (n[:70], c["_total"]) + " | ".join(str(c[x]) if c[x] else "" for x in cols) + " |")
tot = collections.Counter()
for c in kern.values():
    tot.update(c)
print("\nTotals over %d kernels: " % len(kern) + ", ".join("%s %d" % (x, tot[x]) for x in cols if tot[x]))
