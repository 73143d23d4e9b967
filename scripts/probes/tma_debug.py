import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rescan_line_sted_b200 import _lib
lib = _lib.get()
rng = np.random.default_rng(0)
shape = tuple(int(v) for v in sys.argv[1:3]) if len(sys.argv) > 2 else (8, 2048)
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
psfs = rng.random((K, 3, 21))
obj = rng.random((1,) + shape) + 0.1
res = {}
for tma in (0, 1):
    h = _lib.DeconvHandle(lib, psfs, shape, precision=32)
    h.set_option('row_tma', tma)
    h.create_data(obj, 1e6 * obj.size, 1)
    for it in range(2):
        h.iterate(1)
        e = h.get(_lib.ESTIMATE)
        res[(tma, it)] = e
        print('tma', tma, 'iter', it, 'min %.4g max %.4g mean %.6g zeros %.3f nan %d' % (e.min(), e.max(), e.mean(), (e == 0).mean(), np.isnan(e).sum()))
    h.close()
for it in range(2):
    a, b = res[(0, it)], res[(1, it)]
    d = np.abs(a - b)
    print('iter', it, 'rel diff', np.linalg.norm(a - b) / np.linalg.norm(a), 'rows with diff', np.nonzero(d.max(axis=2)[0] > 1e-3 * a.max())[0][:16],
          'cols with diff (first)', np.nonzero(d.max(axis=1)[0] > 1e-3 * a.max())[0][:12])
