#!/usr/bin/env python3
"""Per-kernel SASS instruction summary of liblsted.so (cuobjdump -sass, sm_100a): the
mnemonics that prove how data moves (TMA tensor copies UTMALDG/UTMASTG, bulk copies UBLKCP,
L2 prefetches UBLKPF/UTMAPF/CCTL, mbarrier SYNCS, LDGSTS), packed fp32 (FFMA2/FADD2/FMUL2),
fp64 (DFMA/DADD/DMUL), shared-memory traffic (LDS/STS), barriers (BAR) and local-memory
spills (LDL/STL).  usage: sass_summary.py [lib] > profiles/rNN_sass_summary.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'rescan_line_sted_b200/liblsted.so'
GROUPS = [('UTMALDG', 'UTMALDG'), ('UTMASTG', 'UTMASTG'), ('UTMAPF', 'UTMAPF'), ('UBLKCP', 'UBLKCP'),
          ('UBLKPF', 'UBLKPF'), ('SYNCS', 'SYNCS'), ('LDGSTS', 'LDGSTS'), ('LDG', 'LDG'), ('STG', 'STG'),
          ('LDS', 'LDS'), ('STS', 'STS'), ('BAR', 'BAR'), ('FFMA2', 'FFMA2'), ('FADD2', 'FADD2'),
          ('FMUL2', 'FMUL2'), ('FFMA', 'FFMA'), ('DFMA', 'DFMA'), ('DADD', 'DADD'), ('DMUL', 'DMUL'),
          ('MUFU', 'MUFU'), ('LDL', 'LDL'), ('STL', 'STL')]
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = {}
names = re.findall(r'Function : (\S+)', out)
dm = subprocess.run(['cu++filt'] + names, capture_output=True, text=True).stdout.splitlines()
for n, d in zip(names, dm):
    demangle[n] = d
cur, counts, total = None, {}, collections.Counter()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1)
        counts[cur]['total'] += 1
        for key, prefix in GROUPS:
            if op == prefix or (key in ('LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL', 'BAR', 'MUFU', 'SYNCS') and op.startswith(prefix) and not op.startswith('LDGSTS')):
                counts[cur][key] += 1
                break
            if key not in ('LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL', 'BAR', 'MUFU', 'SYNCS', 'FFMA') and op.startswith(prefix):
                counts[cur][key] += 1
                break


def short(name):
    d = demangle.get(name, name)
    d = re.sub(r'^void ', '', d)
    d = re.sub(r'lsted::', '', d)
    return d if len(d) <= 110 else d[:107] + '...'


keys = [k for k, _ in GROUPS]
print('# SASS instruction summary of %s (cuobjdump -sass, %d kernels, sm_100a)\n' % (lib, len(counts)))
print('Static instruction counts per kernel; columns with only zeros are left out of a row.\n')
print('| kernel | instructions | ' + ' | '.join(keys) + ' |')
print('|---|---|' + '---|' * len(keys))
for name in sorted(counts, key=lambda n: -counts[n]['total']):
    c = counts[name]
    print('| `%s` | %d | ' % (short(name), c['total']) + ' | '.join(str(c[k]) if c[k] else '' for k in keys) + ' |')
    total.update(c)
print('\n**Whole library:** ' + ', '.join('%s %d' % (k, total[k]) for k in keys if total[k]))
