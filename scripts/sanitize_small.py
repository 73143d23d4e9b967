#!/usr/bin/env python3
"""Small shapes that send every fast-path kernel through compute-sanitizer in a minute:
  (8, 2048)   -> 2160-wide fast ROW kernels (all five modes), generic columns
  (2100, 8)   -> 2160-long fast COLUMN kernels (bulk-copy OTF staging, mbarrier), generic rows
  (4, 2048) with option row_dual -> the two-pair ROW_MID kernel
plus the PSF / rotate kernels.  usage (on the GPU box):
  compute-sanitizer --tool racecheck python scripts/sanitize_small.py
  compute-sanitizer --tool memcheck  python scripts/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rescan_line_sted_b200 import _lib, line_sted_tools as st, orientations  # noqa: E402

rng = np.random.default_rng(0)
lib = _lib.get()
for shape, pshape, opts in (((8, 2048), (3, 21), {}), ((2100, 8), (21, 3), {}),
                            ((4, 2048), (3, 9), {'row_dual': 1})):
    for precision in (32, 64):
        psfs = rng.random((2,) + pshape)
        h = _lib.DeconvHandle(lib, psfs, shape, precision=precision)
        for k, v in opts.items():
            h.set_option(k, v)
        x = rng.random((1,) + shape) + 0.1
        h.create_data(x, 1e6 * x.size, 1)          # ROW_FWD, COL_H, ROW_INV_SIM (Poisson queue)
        h.iterate(2)                                # COL_H, ROW_MID, COL_HT, ROW_FINAL (+ norm)
        h.Ht(rng.random((2,) + shape), True)        # ROW_INV_STORE
        est = h.get(_lib.ESTIMATE)
        assert np.isfinite(est).all()
        info = h.info()
        print(shape, precision, opts, 'L =', (info.Ly, info.Lx), 'estimate mean %.6g' % est.mean())
        h.close()
rep = st.psf_report('line', 1, 9, 8, 1, verbose=False)
rot = orientations.rotate_many(rep['psfs']['rescan_sted'], [0, 30.0, 90, 135.0])
print('psf_report + rotate ok', [float(r.sum()) for r in rot][:2])
