#!/usr/bin/env python
"""Top source lines / SASS instructions by stall samples from an ncu source-page CSV
(ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > X.csv)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
mode = sys.argv[2] if len(sys.argv) > 2 else "lines"
cur_file = None
agg = collections.OrderedDict()
sass = []
cur_line = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) >= 9 and r[0].isdigit():
        try:
            samples, inst, tinst = int(r[4]), int(r[7]), int(r[8])
        except ValueError:
            continue
        k = (cur_file, int(r[0]))
        cur_line = k
        a = agg.setdefault(k, [0, 0, 0, r[1].strip()[:84]])
        a[0] += samples
        a[1] += inst
        a[2] += tinst
    elif len(r) >= 9 and r[0] == '' and r[2].startswith('0x'):
        try:
            sass.append((int(r[4]), int(r[7]), cur_line, r[3].strip()[:70]))
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
if mode == "lines":
    for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
        print("%5.1f%% samp %5.1f%% inst lanes %4.1f  %-16s:%4d  %s" % (
            100 * a[0] / tot, 100 * a[1] / toti, a[2] / max(a[1], 1), k[0][:16], k[1], a[3]))
else:
    for s in sorted(sass, key=lambda x: -x[0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
        print("%5.1f%% samp  inst %9d  %-22s %s" % (100 * s[0] / tot, s[1], "%s:%d" % (s[2][0][:14], s[2][1]), s[3]))
