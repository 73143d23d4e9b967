"""Config 3: 64-point line-STED sweep (8 excitation x 8 depletion brightnesses) and
one tune_psf, timed through the public API (host Gaussian fits included)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rescan_line_sted_b200 import line_sted_tools as st
exc = [0.05, 0.1, 0.25, 0.5, 1, 2, 4, 8]
dep = [0, 1, 3, 9, 27, 54, 81, 108]
E, D = np.meshgrid(exc, dep, indexing='ij')
out = {}
for steps in (8, 25):
    st.psf_report_batch('line', E.ravel()[:2], D.ravel()[:2], steps, 1)   # warm-up
    t0 = time.perf_counter(); reps = st.psf_report_batch('line', E.ravel(), D.ravel(), steps, 1)
    t_batch = time.perf_counter() - t0
    t0 = time.perf_counter()
    for e, d in zip(E.ravel()[:16], D.ravel()[:16]):
        st.psf_report('line', e, d, steps, 1, verbose=False)
    t_single = (time.perf_counter() - t0) / 16
    t0 = time.perf_counter()
    for r in reps[:16]:
        for k in ('excitation', 'sted', 'rescan_sted'):
            st.get_width(r['psfs'][k][0, r['psfs'][k].shape[1] // 2, :])
    t_fit = (time.perf_counter() - t0) / 16
    out['steps_%d' % steps] = {'n': reps[0]['psfs']['sted'].shape[-1], 'batch64_s': t_batch,
                               'points_per_s_batch': 64 / t_batch, 'single_call_ms': t_single * 1e3,
                               'host_fits_per_report_ms': t_fit * 1e3}
t0 = time.perf_counter()
res = st.tune_psf('line', 'rescanned', 4.07614, 3.0227, max_excitation_brightness=0.25,
                  steps_per_improved_psf_width=4.)
out['tune_psf_line_rescanned_R4.08_s'] = time.perf_counter() - t0
print(json.dumps(out))
