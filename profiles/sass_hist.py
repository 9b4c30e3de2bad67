"""Opcode histogram / hot regions of an `ncu --page source --csv --print-source sass` export (gzip ok).
   python profiles/sass_hist.py <file.csv[.gz]> <units (e.g. readouts in the launch)> [regions]"""
import collections
import csv
import gzip
import re
import sys

path, units = sys.argv[1], float(sys.argv[2])
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rd = csv.reader(f)
next(rd)
hdr = next(rd)
idx = {h: i for i, h in enumerate(hdr)}
ci, si, ss = idx["Instructions Executed"], idx["Source"], idx["# Samples"]
rows = []
for r in rd:
    try:
        rows.append((r[0], r[si].strip(), int(float(r[ci])), int(float(r[ss]))))
    except (ValueError, IndexError):
        pass
tot = sum(r[2] for r in rows)
samp = sum(r[3] for r in rows)
print("SASS lines %d (%.0f KB), executed %d = %.1f per unit, samples %d" % (len(rows), len(rows) * 16 / 1024, tot, tot / units, samp))
byop, bys = collections.Counter(), collections.Counter()
for a, s, c, sm in rows:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
    op = m.group(2) if m else "?"
    byop[op] += c
    bys[op] += sm
print("opcode: executed per unit | share of stall samples")
for op, c in byop.most_common(36):
    print("  %-10s %8.1f  %5.1f%%" % (op, c / units, 100.0 * bys[op] / max(1, samp)))
live = sum(1 for r in rows if r[2] > 0)
print("lines executed at least once: %d (%.0f KB)" % (live, live * 16 / 1024))
hot = sum(1 for r in rows if r[2] >= 0.2 * units)
print("lines executed >= 0.2 per unit: %d (%.0f KB)" % (hot, hot * 16 / 1024))
if len(sys.argv) > 3:
    # contiguous regions by execution-count level
    start, acc, accs, last = 0, 0, 0, None
    for i, (a, s, c, sm) in enumerate(rows + [("", "", -1, 0)]):
        lvl = -1 if c < 0 else (0 if c == 0 else int(round(4 * (c / units) ** 0.5)))
        if lvl != last:
            if last is not None and i > start and acc / units > 5:
                print("  lines %5d-%5d  %7.1f per unit  samples %4.1f%%  first: %s" % (start, i - 1, acc / units, 100.0 * accs / max(1, samp), rows[start][1][:60]))
            start, acc, accs, last = i, 0, 0, lvl
        acc += max(c, 0)
        accs += sm
