"""Per-frame device times (ms) of the bench workload, for diagnosing jitter."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from rescan_line_sted_b200 import _lib, line_sted_tools as st
N, K, n_iter = 2048, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 64
base = st.psf_report('line', verbose=False, **bench.FIG2_2P0X_LR)['psfs']['rescan_sted']
psfs = bench.orientation_psfs(base, K)
h = _lib.DeconvHandle(_lib.get(), st._stack_psfs(psfs), (N, N), precision=32)
obj = bench.synthetic_object(N)
h.upload_object(obj)
for mode in ('async', 'sync-each'):
    ts = []
    for i in range(8):
        h.sync(); t0 = time.perf_counter()
        h.timer_start()
        h.set_option('forget_normalization', 1)
        h.simulate(bench.total_brightness(N), i)
        t1 = time.perf_counter()
        h.iterate(n_iter)
        t2 = time.perf_counter()
        ms = h.timer_stop()
        ts.append((round(ms, 2), round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2)))
    print(mode, ts)
h.timer_start()
for i in range(5):
    h.set_option('forget_normalization', 1); h.simulate(bench.total_brightness(N), i); h.iterate(n_iter)
print('5 frames back to back: %.2f ms/frame' % (h.timer_stop() / 5))
def run(label, fn, n=5):
    h.sync(); h.timer_start()
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    cpu = (time.perf_counter() - t0) * 1e3 / n
    print('%-46s %.2f ms/frame (cpu enqueue %.2f)' % (label, h.timer_stop() / n, cpu))
B = bench.total_brightness(N)
run('iterate(64) only', lambda i: h.iterate(n_iter))
run('simulate only', lambda i: h.simulate(B, i))
run('simulate+iterate (norm kept)', lambda i: (h.simulate(B, i), h.iterate(n_iter)))
run('forget+simulate+iterate', lambda i: (h.set_option('forget_normalization', 1), h.simulate(B, i), h.iterate(n_iter)))
run('forget+simulate+iterate+sync', lambda i: (h.set_option('forget_normalization', 1), h.simulate(B, i), h.iterate(n_iter), h.sync()))
run('iterate(64) only again', lambda i: h.iterate(n_iter))
