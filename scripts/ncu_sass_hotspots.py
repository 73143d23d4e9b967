#!/usr/bin/env python3
"""Stall-sample breakdown of one kernel from an .ncu-rep SASS page:
totals per stall reason and the hottest instructions.
usage: ncu_sass_hotspots.py rep launch_index [top_n]"""
import csv, subprocess, sys, collections
rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass',
                      '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1][:120])
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data, seen = [], set()
for r in rows[2:]:      # (ncu prints the listing twice when the report holds source + SASS)
    if len(r) == len(hdr) and r[0] != 'Address' and r[0] not in seen:
        seen.add(r[0]); data.append(r)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter()
for r in data:
    for s in stalls:
        try: tot[s] += float(r[idx[s]])
        except ValueError: pass
total = sum(tot.values())
print('total samples', total, ' instructions', len(data))
print('  '.join('%s %.1f%%' % (k[6:], 100 * v / total) for k, v in tot.most_common(9)))
ex = sum(float(r[idx['Instructions Executed']] or 0) for r in data)
print('warp instructions executed', ex)
# opcode histogram by executed count
ops = collections.Counter()
for r in data:
    op = r[idx['Source']].split()[0] if r[idx['Source']].split() else '?'
    if op.startswith('@'): op = r[idx['Source']].split()[1]
    ops[op.split('.')[0]] += float(r[idx['Instructions Executed']] or 0)
print('opcodes:', '  '.join('%s %.1f%%' % (k, 100 * v / ex) for k, v in ops.most_common(18)))
hot = sorted(data, key=lambda r: -float(r[idx['# Samples']] or 0))[:top]
for r in hot:
    s = {k[6:]: float(r[idx[k]] or 0) for k in stalls}
    main = max(s, key=s.get)
    print('%6s %5.2f%% %-60s %s' % (r[idx['Address']][-5:], 100 * float(r[idx['# Samples']]) / total,
                                   r[idx['Source']][:60], main))
