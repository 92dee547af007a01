#!/usr/bin/env python
"""Stall-reason totals and the hottest SASS lines of one kernel in an .ncu-rep (source page).
usage: python tools/ncu_stalls.py rep [kernel-substring] [top-N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Kernel Name"') and pat in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
tab = list(csv.reader(io.StringIO("\n".join(lines[start + 1:end]))))
h = tab[0]
ci, cs, cn = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot, rows, ninst = {}, [], 0
for idx, r in enumerate(tab[1:]):
    if len(r) < len(h):
        continue
    try:
        s = int(r[cs])
    except ValueError:
        continue
    ninst += int(r[cn]) if r[cn].isdigit() else 0
    rows.append((s, idx, r))
    for i in stall_cols:
        if r[i].isdigit():
            tot[h[i]] = tot.get(h[i], 0) + int(r[i])
T = sum(s for s, _, _ in rows)
print("samples %d, warp instructions %d" % (T, ninst))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
    print("  %-24s %9d %.3f" % (k, v, v / T))
for s, idx, r in sorted(rows, key=lambda x: -x[0])[:topn]:
    top = sorted(((int(r[i]) if r[i].isdigit() else 0, h[i][6:]) for i in stall_cols), reverse=True)[:2]
    print("%7d %5.2f%% #%-5d %-58s %s" % (s, 100 * s / T, idx, r[ci][:58], top))
