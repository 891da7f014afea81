#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --kernel-id :::N` output: stall reasons in total
and the hottest SASS lines.  usage: ncu_stalls.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]
ia, isamp = hdr.index("Source"), hdr.index("# Samples")
stalls = [i for i, name in enumerate(hdr) if name.startswith("stall_") and "Not Issued" not in name]
tot, lines = {}, []
for r in rows[h + 1:]:
    if len(r) < len(hdr) or not r[isamp].isdigit():
        continue
    n = int(r[isamp])
    if n == 0:
        continue
    st = {hdr[i][6:]: int(r[i] or 0) for i in stalls if (r[i] or "0").isdigit() and int(r[i] or 0) > 0}
    lines.append((n, r[ia].strip(), st))
    for k, v in st.items():
        tot[k] = tot.get(k, 0) + v
s = sum(tot.values())
print("kernel:", rows[0][1][:100] if rows and len(rows[0]) > 1 else "?")
print("samples", s, "by reason:", [(k, round(100 * v / s, 1)) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])])
for n, src, st in sorted(lines, key=lambda t: -t[0])[:top]:
    print(f"{100*n/s:5.1f}%  {src[:64]:64s} {st}")
