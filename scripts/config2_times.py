#!/usr/bin/env python3
"""Config 2 (BASELINE.json configs[1]; line_sted_figure_2.py:33-57): Richardson-Lucy
iterations/s at the reference's own sizes -- 128^2 object, 107^2 PSFs, K in {1, 4, 10}
orientations -- through the drop-in `Deconvolver.iterate()` loop the figure script runs,
for one deconvolver and for the 24 concurrent deconvolvers of figure 2 (each its own handle
and CUDA stream, iterated round-robin exactly like :53-56).  The CPU column is the unmodified
reference (baseline/_ref) when present, else the oracle port, on a bounded iteration count.

    python scripts/config2_times.py [--out profiles/r02_config2.json] [--iterations 256]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def make_psfs(K):
    """K orientations of the figure-2 '2p0x_lr' rescan PSF (107^2), through the product API."""
    from rescan_line_sted_b200 import line_sted_tools as st, orientations
    rep = st.psf_report('line', 0.21672180512595912, 11.766131198861775, 25, 5, verbose=False)
    return orientations.line_orientation_psfs(rep['psfs']['rescan_sted'], K, 3.0227)


def test_object(N=128):
    y, x = np.mgrid[0:N, 0:N]
    return (100 + 80 * np.cos(x / 4.0) * np.sin(y / 7.0) + 40 * ((x // 16 + y // 16) % 2))[None].astype(np.float64)


def time_gpu(st, psfs, obj, deconvolvers, iterations, precision):
    os.environ['LSTED_PRECISION'] = precision
    ds = [st.Deconvolver(psfs, output_prefix='/tmp/lsted_c2_%d_' % i, verbose=False)
          for i in range(deconvolvers)]
    for i, d in enumerate(ds):
        d.create_data_from_object(obj, total_brightness=5e10, random_seed=i)
    for d in ds:                    # warm-up (normalisation, first launches)
        for _ in range(4):
            d.iterate()
        d.estimate
    t = time.perf_counter()
    for _ in range(iterations):     # line_sted_figure_2.py:53-56
        for d in ds:
            d.iterate()
    ests = [d.estimate for d in ds]
    dt = time.perf_counter() - t
    assert all(np.isfinite(e).all() and e.max() > 0 for e in ests)
    return dt / (iterations * deconvolvers)


def time_cpu(psfs, obj, iterations):
    try:
        from _reference_loader import load_reference, reference_available
        if not reference_available():
            raise ImportError
        mod, kind = load_reference(), 'reference'
    except Exception:
        from oracle import line_sted_oracle as mod
        kind = 'port'
    d = mod.Deconvolver(psfs, verbose=False) if kind == 'reference' else mod.Deconvolver(psfs)
    d.create_data_from_object(obj, total_brightness=5e10, random_seed=0)
    d.iterate()
    t = time.perf_counter()
    for _ in range(iterations):
        d.iterate()
    return (time.perf_counter() - t) / iterations, kind


def bench_record(iterations=128):
    """Short form for bench.py's `extra.config2`: K = 4, 24 concurrent deconvolvers."""
    from rescan_line_sted_b200 import line_sted_tools as st
    obj = test_object()
    psfs = make_psfs(4)
    prev = os.environ.get('LSTED_PRECISION')
    rec = {'object': 128, 'psf': 107, 'K': 4, 'deconvolvers': 24,
           'api': 'line_sted_tools.Deconvolver.iterate() round-robin over 24 deconvolvers '
                  '(line_sted_figure_2.py:53-56), wall clock incl. the final read of every estimate'}
    try:
        for precision in ('fp32', 'fp64'):
            s = time_gpu(st, psfs, obj, 24, max(iterations // 8, 8), precision)
            rec['%s_rl_iterations_per_s' % precision] = 1.0 / s
            rec['%s_us_per_iteration' % precision] = s * 1e6
    finally:
        if prev is None:
            os.environ.pop('LSTED_PRECISION', None)
        else:
            os.environ['LSTED_PRECISION'] = prev
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--iterations', type=int, default=256)
    ap.add_argument('--cpu-iterations', type=int, default=10)
    args = ap.parse_args()
    from rescan_line_sted_b200 import line_sted_tools as st
    obj = test_object()
    rows = []
    for K in (1, 4, 10):
        psfs = make_psfs(K)
        cpu_s, kind = time_cpu(psfs, obj, args.cpu_iterations)
        row = dict(K=K, object=128, psf=107, cpu_ms_per_iteration=cpu_s * 1e3, cpu_kind=kind,
                   cpu_iterations_timed=args.cpu_iterations)
        for precision in ('fp32', 'fp64'):
            for nd in (1, 24):
                s = time_gpu(st, psfs, obj, nd, args.iterations if nd == 1 else max(args.iterations // 8, 8),
                             precision)
                row['%s_%d_deconvolvers_us_per_iteration' % (precision, nd)] = s * 1e6
                row['%s_%d_deconvolvers_iterations_per_s' % (precision, nd)] = 1.0 / s
        row['speedup_fp64_24_vs_cpu'] = row['cpu_ms_per_iteration'] * 1e-3 * row['fp64_24_deconvolvers_iterations_per_s']
        rows.append(row)
        print(json.dumps(row))
    # per-kernel device times of one deconvolver (K = 4), CUDA events on the handle's stream
    kernels = {}
    for precision in ('fp32', 'fp64'):
        os.environ['LSTED_PRECISION'] = precision
        d = st.Deconvolver(make_psfs(4), output_prefix='/tmp/lsted_c2p_', verbose=False)
        d.create_data_from_object(obj, total_brightness=5e10, random_seed=0)
        for _ in range(4):
            d.iterate()
        h = d._handle
        h.set_option('profile', 1)
        h.profile(reset=True)
        for _ in range(64):
            d.iterate()
        d.estimate
        kernels[precision] = {k: dict(avg_us=1e3 * ms / n, launches=n) for k, (ms, n) in h.profile().items() if n}
        h.set_option('profile', 0)
    print(json.dumps(kernels))
    if args.out:
        with open(args.out, 'w') as f:
            json.dump(dict(rows=rows, cpu_cores=os.cpu_count(), kernels_K4=kernels), f, indent=1)


if __name__ == '__main__':
    main()
