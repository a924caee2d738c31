#!/usr/bin/env python
"""Per-CUDA-source-line summary of an `ncu --page source --print-source cuda,sass --csv` dump:
share of executed instructions, share of stall samples, average active threads."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]
iS, iI, iT = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
lines = []
for r in rows[hi + 1:]:
    if r[0] != '':
        try:
            lines.append((int(r[0]), r[1].strip()[:100], int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0)))
        except ValueError:
            pass
totI = sum(l[3] for l in lines)
totS = sum(l[2] for l in lines)
totT = sum(l[4] for l in lines)
print(f"total warp-inst {totI}  thread-inst {totT}  avg active threads {totT / totI:.2f}  stall samples {totS}")
print("%5s %8s %7s %8s  %s" % ("line", "inst%", "stall%", "thr/inst", "source"))
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    print("%5d %7.2f%% %6.2f%% %8.1f  %s" % (l[0], 100 * l[3] / totI, 100 * l[2] / totS, l[4] / max(l[3], 1), l[1]))
