#!/usr/bin/env python3
"""Throughput of the figure-3 scan-position engine on one GPU, next to the
CPU restatement of the reference loop (oracle/scan_oracle.py, the reference's own scipy calls)
timed on a bounded sample of scan positions on this box's host.

    python scripts/scan_times.py [--out profiles/r02_scan.json] [--big]

Workloads: the figure-3 script's own cases (line_sted_figure_3.py:40-63: 128^2 objects, 1x1 and
2x2 fields of view, R = 1..3, psf_width 25) and, with --big, a 1024^2 field of view.
Reported per case: scan positions, device ms per orientation (CUDA events inside liblsted),
scan positions / s, and the algorithmic bytes rate: per scan position the engine must read the
object plane and write the detector plane once (2 * 8 * n0 * n1 bytes, fp64)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_object(n):
    """Stripe / ring texture in [1e-6, 1] like the reference's 8-bit test objects."""
    y, x = np.mgrid[0:n, 0:n]
    r = np.hypot(y - n / 2, x - n / 2)
    return (0.5 + 0.5 * np.cos(r / 3.0) * np.cos(x / 5.0))[None] * (1 - 1e-6) + 1e-6


def cpu_seconds_per_position(obj, typ, width, R, pad, sample):
    """The reference loop body (oracle: same scipy calls) on `sample` positions of rot 0."""
    from oracle import scan_oracle as so
    from scipy.ndimage import gaussian_filter
    psf_sigma = width / (2 * np.sqrt(2 * np.log(2)))
    step, positions, exc_sep = so.scan_plan(obj.shape, typ, width, R, pad)
    o = np.pad(obj, ((0, 0), (pad, pad), (pad, pad)), 'constant')
    cexc = np.zeros(o.shape)
    cexc[0, o.shape[1] // 2, :] = 1
    cexc = gaussian_filter(cexc, (0, psf_sigma / R, 0), truncate=8)
    pick = positions[::max(1, len(positions) // sample)][:sample]
    t = time.perf_counter()
    for (sy, sx) in pick:
        exc = so.shift(cexc, (0, sy, sx))
        glow = o * exc
        desc = so.shift(glow, (0, -sy, -sx))
        so.rotate(exc, -30.0); so.rotate(glow, -30.0); so.rotate(desc, -30.0)   # :169-171
        if typ == 'rescan_line':
            so.shift(so.scale_y(gaussian_filter(desc, psf_sigma), 1 / (R ** 2 + 1)), (0, sy, sx))
        else:
            gaussian_filter(desc if typ != 'nondescan_multipoint' else glow, psf_sigma)
    return (time.perf_counter() - t) / len(pick), len(pick)


def bench_record():
    """Short form for bench.py's `extra.scan_engine`: the figure-3 rescan-line and descan-point
    cases at the script's largest size (2x2 field of view, R = 3)."""
    from rescan_line_sted_b200 import scan_engine as se
    rec = {}
    for typ, n_or, pad in (('rescan_line', 6, int(0.45 * 256)), ('descan_point', 1, 25)):
        obj = synthetic_object(256)
        se.simulate_imaging(obj, typ, 25, 3, 1, 1, pad, verbose=False)      # warm-up
        out = se.simulate_imaging(obj, typ, 25, 3, n_or, 1, pad, verbose=False)
        P = len(out['scan_positions'])
        runs = (n_or if typ.endswith('line') else 1) + 1
        n0 = 256 + 2 * pad
        rec[typ] = {'object': 256, 'padded': n0, 'R': 3, 'orientations': n_or, 'scan_positions': P,
                    'device_ms': out['device_ms'],
                    'positions_per_s': P * runs / (out['device_ms'] * 1e-3),
                    'algorithmic_GBps': 2 * 8 * n0 * n0 * P * runs / (out['device_ms'] * 1e-3) / 1e9}
    rec['note'] = ('simulate_imaging (line_sted_figure_3.py:76-273) on the device, all scan positions of '
                   'an orientation in flight; device_ms = CUDA events inside liblsted over the scan of '
                   'every orientation + the display-maxima pass')
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--big', action='store_true')
    ap.add_argument('--cpu-sample', type=int, default=3)
    args = ap.parse_args()
    from rescan_line_sted_b200 import scan_engine as se
    cases = []
    for fov in (1, 2):
        for R in (1, 2, 3):
            for typ in ('descan_point', 'nondescan_multipoint'):
                cases.append((128 * fov, typ, 25, R, 1, 25))
            for typ in ('descan_line', 'rescan_line'):
                cases.append((128 * fov, typ, 25, R, 6, int(0.45 * 128 * fov)))
    if args.big:
        for typ in ('descan_line', 'rescan_line'):
            cases.append((1024, typ, 25, 3, 6, int(0.45 * 1024)))
    rows = []
    for n, typ, width, R, n_or, pad in cases:
        obj = synthetic_object(n)
        se.simulate_imaging(obj, typ, width, R, min(n_or, 2), 1, pad, verbose=False)   # warm-up
        t = time.perf_counter()
        out = se.simulate_imaging(obj, typ, width, R, n_or, 1, pad, verbose=False)
        wall = time.perf_counter() - t
        P = len(out['scan_positions'])
        runs = (n_or if typ.endswith('line') else 1) + 1          # + the find_maxima pass
        n0 = n + 2 * pad
        row = dict(object=n, padded=n0, imaging_type=typ, R=R, orientations=n_or, scan_positions=P,
                   device_ms_per_orientation=out['device_ms'] / runs,
                   positions_per_s=P * runs / (out['device_ms'] * 1e-3),
                   algorithmic_GBps=2 * 8 * n0 * n0 * P * runs / (out['device_ms'] * 1e-3) / 1e9,
                   wall_s=wall)
        if typ in ('descan_line', 'rescan_line') and args.cpu_sample and n <= 256:
            sec, cnt = cpu_seconds_per_position(obj, typ, width, R, pad, args.cpu_sample)
            row['cpu_s_per_position'] = sec
            row['cpu_sample_positions'] = cnt
            row['speedup_vs_cpu_loop'] = sec * P * runs / (out['device_ms'] * 1e-3)
        rows.append(row)
        print(json.dumps(row))
    if args.out:
        with open(args.out, 'w') as f:
            json.dump(dict(cpu_cores_used=1, rows=rows), f, indent=1)


if __name__ == '__main__':
    main()
