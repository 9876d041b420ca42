#!/usr/bin/env python
"""Join an ncu SASS source-page CSV with nvdisasm line info -> instructions executed per CUDA source line.

usage: line_profile.py <ncu --page source --csv output> <nvdisasm --print-line-info output> <mangled kernel substring>
"""
import collections
import csv
import re
import sys

src_csv, dis, kern = sys.argv[1:4]
# --- nvdisasm: ordered list of (line-tag) per instruction of the kernel
tags, cur, on = [], None, False
for ln in open(dis, errors="replace"):
    if ln.startswith("\t.section") or ln.startswith("//---"):
        on = kern in ln and ".text." in ln if ".text." in ln else on
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln):
        tags.append(cur)
rows = list(csv.reader(open(src_csv)))
H = rows[1]
ci, ti, si, sm = H.index("Instructions Executed"), H.index("Thread Instructions Executed"), H.index("Source"), H.index("# Samples")
ins = [(int(r[ci]), int(r[ti]), int(r[sm]), r[si]) for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
print("sass instrs: ncu %d, nvdisasm %d" % (len(ins), len(tags)))
agg = collections.defaultdict(lambda: [0, 0, 0])
for (n, t, s, _), tag in zip(ins, tags):
    a = agg[tag]
    a[0] += n; a[1] += t; a[2] += s
tot = sum(a[0] for a in agg.values()); stot = sum(a[2] for a in agg.values())
print("total warp instr %d, samples %d" % (tot, stot))
for tag, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 50]:
    print("%5.1f%% inst  %5.1f%% samples  lanes %4.1f  %s" % (100 * a[0] / tot, 100 * a[2] / max(stot, 1), a[1] / max(a[0], 1), tag))

if len(sys.argv) > 5:  # phase ranges "name:file:lo-hi,..."
    print("--- phases")
    for spec in sys.argv[5].split(","):
        name, f, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        n = sum(a[0] for t, a in agg.items() if t and t[0] == f and lo <= t[1] <= hi)
        th = sum(a[1] for t, a in agg.items() if t and t[0] == f and lo <= t[1] <= hi)
        s = sum(a[2] for t, a in agg.items() if t and t[0] == f and lo <= t[1] <= hi)
        print("%-28s %5.1f%% inst  %5.1f%% samples  lanes %4.1f   (%.1f warp-inst per 1000 total)" % (name, 100 * n / tot, 100 * s / max(stot, 1), th / max(n, 1), 1000 * n / tot))
