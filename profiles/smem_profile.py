#!/usr/bin/env python
"""Per CUDA source line: shared-memory wavefronts (actual / ideal), instructions, stall samples by reason.

usage: smem_profile.py <ncu --page source --csv> <nvdisasm --print-line-info output> <kernel substring> [top N]
(same join as line_profile.py: the n-th SASS instruction of the ncu page is the n-th instruction nvdisasm lists)
"""
import collections
import csv
import re
import sys

src_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tags, cur, on = [], None, False
for ln in open(dis, errors="replace"):
    if ln.startswith("\t.section") or ln.startswith("//---"):
        on = kern in ln and ".text." in ln if ".text." in ln else on
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln):
        tags.append(cur)
rows = list(csv.reader(open(src_csv)))
H = rows[1]
col = {n: H.index(n) for n in ("Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "Source",
                               "stall_barrier", "stall_short_sb", "stall_long_sb", "stall_wait", "stall_no_inst", "stall_mio", "stall_math", "stall_branch_resolving")}
data = [r for r in rows[2:] if len(r) > col["Instructions Executed"] and r[col["Instructions Executed"]].isdigit()]
assert len(data) == len(tags), (len(data), len(tags))
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r, tag in zip(data, tags):
    for k, c in col.items():
        if k == "Source":
            continue
        v = int(r[c]) if r[c].isdigit() else 0
        agg[tag][k] += v
        tot[k] += v
print("totals:", dict(tot))
def show(key, title):
    print("--- top lines by", title)
    for tag, a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
        print("%6.2f%%  inst %5.2f%%  wavefronts %9d (ideal %9d)  samples %5d  bar %4d ssb %4d lsb %4d wait %4d noinst %4d  %s" % (
            100.0 * a[key] / max(tot[key], 1), 100.0 * a["Instructions Executed"] / tot["Instructions Executed"], a["L1 Wavefronts Shared"], a["L1 Wavefronts Shared Ideal"],
            a["# Samples"], a["stall_barrier"], a["stall_short_sb"], a["stall_long_sb"], a["stall_wait"], a["stall_no_inst"], tag))
show("L1 Wavefronts Shared", "shared-memory wavefronts")
show("# Samples", "stall samples")
