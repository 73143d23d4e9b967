#!/usr/bin/env python3
"""Compare each kernel's share of the timed step as bench.py reports it (CUDA events,
`kernels`) with its share in the ncu launch list of the same command
(`ncu --metrics gpu__time_duration.sum --clock-control none`).  ncu's per-launch times
are cold-cache and serialised, so only the shares are comparable.
usage: launch_share_check.py launches.csv bench.json"""
import collections
import csv
import json
import re
import sys

NAMES = [('row_fast_kernel<0', 'row_fwd'), ('row_fast_kernel<1', 'row_inv_store'),
         ('row_fast_kernel<2', 'row_inv_sim'), ('row_fast_kernel<3', 'row_mid'),
         ('row_mid_dual_kernel', 'row_mid'), ('row_fast_kernel<4', 'row_final'),
         ('col_fast_kernel<1', 'col_h'), ('col_fast_kernel<2', 'col_ht'),
         ('col_sub_kernel<1', 'col_h'), ('col_sub_kernel<2', 'col_ht')]


def main(csv_path, json_path):
    rows = [r for r in csv.reader(l for l in open(csv_path) if not l.startswith('=='))]
    hdr = rows[0]
    name_i, val_i, unit_i = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rows[1:]:
        if len(r) <= val_i:
            continue
        us = float(r[val_i].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[unit_i], 1.0)
        key = 'other'
        for pat, short in NAMES:
            if pat in r[name_i].replace('(int)', ''):
                key = short
        tot[key] += us
        cnt[key] += 1
    bench = json.loads(open(json_path).read().strip().splitlines()[-1])
    ks = bench['kernels']
    all_ncu = sum(tot.values())
    all_ev = sum(v['ms_per_step'] for v in ks.values())
    print('| kernel | ncu launches | ncu share | bench (CUDA events) share |')
    print('|---|---|---|---|')
    for k in ('row_mid', 'col_ht', 'col_h', 'row_final', 'row_inv_sim', 'row_fwd', 'row_inv_store', 'other'):
        ev = ks.get(k, ks.get('elementwise' if k == 'other' else k, {'ms_per_step': 0.0}))['ms_per_step']
        print('| %s | %d | %.3f | %.3f |' % (k, cnt[k], tot[k] / all_ncu, ev / all_ev))


if __name__ == '__main__':
    main(*sys.argv[1:3])
