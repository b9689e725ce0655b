#!/usr/bin/env python
"""Per-source-line executed instructions / stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if not hdr or len(r) < 10 or not r[0]:
        continue                      # SASS rows have an empty line number
    d = dict(zip(hdr, r))
    try:
        agg[(cur, int(r[0]), r[1].strip()[:100])] = (int(d['Instructions Executed']), int(d['# Samples']))
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values()) or 1
smp = sum(v[1] for v in agg.values()) or 1
print('total warp instructions', tot, 'samples', smp)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% inst %6.2f%% smp  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / smp, k[0], k[1], k[2]))
