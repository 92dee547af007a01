#!/usr/bin/env python
"""Samples and instructions between consecutive sync / TMEM marker instructions of a kernel.
usage: python tools/ncu_regions.py rep [kernel-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Kernel Name"') and pat in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
tab = list(csv.reader(io.StringIO("\n".join(lines[start + 1:end]))))
h = tab[0]; ci, cs, cn = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
rows = [r for r in tab[1:] if len(r) > max(ci, cs, cn)]
S = [int(r[cs]) if r[cs].isdigit() else 0 for r in rows]
N = [int(r[cn]) if r[cn].isdigit() else 0 for r in rows]
T = sum(S)
keys = ("UTCBAR", "TRYWAIT", "BAR.SYNC", "LDTM", "UTCHMMA")
marks = [0]
last_kind = None
for i, r in enumerate(rows):
    k = next((k for k in keys if k in r[ci]), None)
    if k and not (k == last_kind and i - marks[-1] < 40):
        marks.append(i); last_kind = k
    elif k:
        last_kind = k
marks.append(len(rows))
print("total samples", T, "instr", sum(N))
for a, b in zip(marks[:-1], marks[1:]):
    s = sum(S[a:b]); n = sum(N[a:b])
    if s * 200 > T or n * 200 > sum(N):
        print("%5d-%5d %-44s samples %8d (%5.1f%%) instr %12d (%4.1f%%) exec %10d" % (a, b, rows[a][ci][:44], s, 100 * s / T, n, 100.0 * n / sum(N), N[a]))
