"""Instructions executed per source line / per function: joins an `ncu --page source --csv --print-source sass` export with the line
table of the cubin (`nvdisasm -g`).   python profiles/sass_by_source.py <sass.csv[.gz]> <nvdisasm -g output> <kernel mangled name> <units>"""
import collections
import csv
import gzip
import os
import re
import sys

sass_csv, disasm, kernel, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.strip().startswith(".section\t.text." + kernel + ","))
ins, cur = {}, (None, None)
for l in lines[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins[int(m.group(1), 16)] = cur
f = gzip.open(sass_csv, "rt") if sass_csv.endswith(".gz") else open(sass_csv)
rd = csv.reader(f)
next(rd)
hdr = next(rd)
idx = {h: i for i, h in enumerate(hdr)}
rows = list(rd)
base = int(rows[0][0], 16)
cnt, smp = collections.Counter(), collections.Counter()
for r in rows:
    key = ins.get(int(r[0], 16) - base, ("?", 0))
    cnt[key] += int(float(r[idx["Instructions Executed"]]))
    smp[key] += int(float(r[idx["# Samples"]]))
tot, ts = sum(cnt.values()), max(1, sum(smp.values()))
print("instructions per unit: %.1f" % (tot / units))
src = {}


def text(fn, ln):
    p = os.path.join(ROOT, "alphago.jl_b200", "csrc", fn)
    if fn not in src:
        src[fn] = open(p).read().split("\n") if os.path.exists(p) else None
    return src[fn][ln - 1].strip()[:90] if src[fn] and 0 < ln <= len(src[fn]) else ""


def functions(fn):
    out = []
    if text(fn, 1) is None or src.get(fn) is None:
        return out
    for i, l in enumerate(src[fn], 1):
        m = re.match(r"\s*(?:template <[^>]*>\s*)?AGZ_(?:DEV|COLD) (?:static )?[\w:<>\*& ]+?\s+(\w+)\(", l)
        if m:
            out.append((i, m.group(1)))
    return out


agg, sagg = collections.Counter(), collections.Counter()
for (fn, ln), c in cnt.items():
    name = "*"
    for i, n in functions(fn) if fn else []:
        if i <= ln:
            name = n
        else:
            break
    agg[(fn, name)] += c
    sagg[(fn, name)] += smp[(fn, ln)]
print("-- by function: instructions per unit | share of stall samples")
for k, c in agg.most_common(24):
    print("%8.1f  %5.1f%%  %s:%s" % (c / units, 100.0 * sagg[k] / ts, k[0], k[1]))
print("-- by line")
for (fn, ln), c in cnt.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 40):
    print("%8.1f  %5.1f%%  %s:%d  %s" % (c / units, 100.0 * smp[(fn, ln)] / ts, fn, ln, text(fn, ln)))
