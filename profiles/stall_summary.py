#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall reasons overall and the hottest SASS lines.
usage: stall_summary.py file.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
lines = []
ninst = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        if r and r[0] == "Kernel Name":
            break                      # next kernel instance in the same export
        continue
    if r[col["# Samples"]] == "# Samples":
        continue
    smp = int(r[col["# Samples"]] or 0)
    ninst += int(r[col["Instructions Executed"]] or 0)
    for s in stalls:
        tot[s] += int(r[col[s]] or 0)
    lines.append((smp, r[col["Source"]].strip(), int(r[col["Instructions Executed"]] or 0),
                  max(stalls, key=lambda s: int(r[col[s]] or 0))))
allsmp = sum(l[0] for l in lines)
print("kernel:", rows[0][1][:100]); print("samples", allsmp, "warp-instructions", ninst)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {s:28s} {100.0 * v / max(allsmp, 1):5.1f}%")
import collections
grp = collections.OrderedDict()
for smp, src, n, why in lines:
    e = grp.setdefault(n, [0, 0, 0, src])
    e[0] += 1; e[1] += n; e[2] += smp
print("instruction groups by execution count (= loop bodies):")
for k, e in sorted(grp.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  count {k:9d} x{e[0]:4d} instrs = {100.0 * e[1] / max(ninst, 1):5.1f}% inst {100.0 * e[2] / max(allsmp, 1):5.1f}% smp   first: {e[3][:50]}")
print("hottest lines:")
for smp, src, n, why in sorted(lines, key=lambda l: -l[0])[:top]:
    print(f"  {100.0 * smp / max(allsmp, 1):5.2f}%  {n:9d}  {why:22s} {src[:80]}")
