#!/usr/bin/env python3
"""Per-source-line executed-instruction counts from an .ncu-rep: python scripts/ncu_lines.py <rep> <file-substr> [top]"""
import csv, subprocess, sys
rep, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
tables, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "hdr": None, "rows": []}; tables.append(cur)
    elif cur is not None and r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(r)
for t in tables:
    if sub not in t["file"]:
        continue
    hdr = t["hdr"]; ie = hdr.index('Instructions Executed'); ln = hdr.index('Line No')
    srcs = [i for i, h in enumerate(hdr) if h == 'Source']
    items, tot = [], 0
    for r in t["rows"]:
        if not r[ln]:
            continue
        try:
            v = int(r[ie].replace(',', ''))
        except ValueError:
            continue
        tot += v; items.append((v, r[ln], r[srcs[0]][:110]))
    items.sort(reverse=True)
    print(t["file"], "warp-instr total", tot)
    for v, l, s in items[:top]:
        print(f"{v:12d} {100 * v / tot:5.1f}% L{l}: {s}")
