#!/usr/bin/env python
"""Hot source lines of one captured kernel launch: aggregates the warp-stall samples of `ncu --page source --print-source cuda,sass`
per CUDA source line (needs -lineinfo and --import-source on).  Usage: ncu_hot_lines.py <report.ncu-rep> [top N] [launch index]"""
import collections
import csv
import subprocess
import sys

src = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict()
cur_file, hdr, launch, first_file = None, None, -1, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        if first_file is None:
            first_file = cur_file
        if cur_file == first_file:
            launch += 1
        continue
    if r[0] == "Line No":
        hdr = r
        idx = {k: hdr.index(k) for k in ("# Samples", "Instructions Executed", "stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_math",
                                        "stall_not_selected", "stall_branch_resolving")}
        continue
    if hdr is None or launch != which or len(r) < 10 or r[2] != "-":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    key = (cur_file.split("/")[-1], ln, r[1].strip()[:100])
    a = agg.setdefault(key, collections.Counter())
    for k, i in idx.items():
        try:
            a[k] += int(r[i])
        except ValueError:
            pass
tot = sum(v["# Samples"] for v in agg.values())
toti = sum(v["Instructions Executed"] for v in agg.values())
print(f"launch {which}: {tot} stall samples, {toti} warp instructions")
byfile = collections.Counter()
for k, v in agg.items():
    byfile[k[0]] += v["# Samples"]
print("samples by file:", [(f, round(100 * n / tot, 1)) for f, n in byfile.most_common(8)])
print("  smp%  inst%   long noinst   wait  short   math notsel branch  where")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    print("%5.1f%% %5.1f%% %6d %6d %6d %6d %6d %6d %6d  %s:%d  %s" % (100 * v["# Samples"] / tot, 100 * v["Instructions Executed"] / max(toti, 1), v["stall_long_sb"],
          v["stall_no_inst"], v["stall_wait"], v["stall_short_sb"], v["stall_math"], v["stall_not_selected"], v["stall_branch_resolving"], k[0], k[1], k[2]))
