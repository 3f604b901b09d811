#!/usr/bin/env python3
"""Hottest stalled SASS instructions (with a few lines of context) of one kernel in an .ncu-rep."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.02
ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
seg = rows[start[0] + 1:(start[1] if len(start) > 1 else None)]
hdr, data = seg[0], seg[1:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
agg = {}
for r in data:
    for k in hdr:
        if k.startswith('stall_') and 'Not' not in k and r[ix[k]].isdigit():
            agg[k] = agg.get(k, 0) + int(r[ix[k]])
print(rows[start[0]][1][:80], 'samples', tot)
print('  '.join('%s=%.1f%%' % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
hot = [i for i, r in enumerate(data) if int(r[ix['# Samples']]) > thr * tot]
for h in hot:
    print('-----')
    for i in range(max(0, h - ctx), h + 1):
        r = data[i]
        st = {k: int(r[ix[k]]) for k in hdr if k.startswith('stall_') and 'Not' not in k and r[ix[k]].isdigit() and int(r[ix[k]]) > 0}
        st = sorted(st.items(), key=lambda x: -x[1])[:2]
        print('%6.2f%% %-72s %s' % (100 * int(r[ix['# Samples']]) / tot, r[ix['Source']][:72], st if i == h else ''))
