#!/usr/bin/env python
"""Top source lines per stall reason from an ncu source-page CSV
(ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > X.csv).
usage: python tools/ncu_stall_lines.py X.csv [reason ...]   (reasons: long_sb, no_inst, wait, ...)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
reasons = sys.argv[2:] or ["long_sb", "no_inst", "wait", "short_sb", "branch_resolving", "barrier"]
hdr = None
cur = None
agg = {r: collections.Counter() for r in reasons}
text = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 40 and r[0] == 'Line No':
        hdr = r
        col = {x: hdr.index('stall_' + x) for x in reasons}
        continue
    if hdr and len(r) > 40 and r[0].isdigit():
        k = (cur, int(r[0]))
        text[k] = r[1].strip()[:80]
        for x in reasons:
            try:
                agg[x][k] += int(r[col[x]])
            except ValueError:
                pass
for x in reasons:
    tot = sum(agg[x].values()) or 1
    print("== stall_%s (%d samples)" % (x, tot))
    for k, v in agg[x].most_common(12):
        print("  %5.1f%%  %s:%d  %s" % (100.0 * v / tot, k[0][:18], k[1], text[k]))
