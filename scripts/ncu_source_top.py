"""Per-kernel digest of `ncu --page source --csv`: executed-instruction mix by opcode and the SASS instructions
that collect the most warp-stall samples (with the dominant stall reason).
  python scripts/ncu_source_top.py <source.csv> [top=25] [kernel-substring]"""
import csv, sys, re
from collections import Counter
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
kernels = []; cur = None; hdr = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; kernels.append(cur); hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; cur["hdr"] = hdr; continue
    if cur is not None and hdr is not None and r: cur["rows"].append(r)
for k in kernels:
    if want and want not in k["name"]: continue
    h = k["hdr"]; iS = h.index("Source"); iN = h.index("# Samples"); iE = h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[iN] or 0) for r in k["rows"]); totE = sum(int(r[iE] or 0) for r in k["rows"])
    print("=== %s\n    samples %d, warp instructions executed %d" % (k["name"][:100], tot, totE))
    mix = Counter(); smix = Counter(); reasons = Counter()
    for r in k["rows"]:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS]); op = m.group(2) if m else "?"
        mix[op] += int(r[iE] or 0); smix[op] += int(r[iN] or 0)
        for i, c in stall_cols: reasons[c] += int(r[i] or 0)
    print("    executed mix: " + ", ".join("%s %.1f%%" % (o, 100.0 * n / totE) for o, n in mix.most_common(16)))
    print("    samples by opcode: " + ", ".join("%s %.1f%%" % (o, 100.0 * n / tot) for o, n in smix.most_common(12)))
    print("    samples by reason: " + ", ".join("%s %.1f%%" % (o[6:], 100.0 * n / tot) for o, n in reasons.most_common(8)))
    order = sorted(range(len(k["rows"])), key=lambda j: -int(k["rows"][j][iN] or 0))[:top]
    for j in sorted(order):
        r = k["rows"][j]; n = int(r[iN] or 0)
        best = max(stall_cols, key=lambda ic: int(r[ic[0]] or 0))
        print("    %5d  %5.1f%%  x%-8s %-14s %s" % (j, 100.0 * n / tot, r[iE], best[1][6:], r[iS].strip()[:90]))
