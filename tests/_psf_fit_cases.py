"""Shared bodies of the device-fit / fused psf_report tests (run on the CPU replay by
tests/test_host_mirror.py and on the GPU by tests/test_gpu_psf.py)."""
import warnings

import numpy as np

SWEEP_EXC = [0.05, 0.1, 0.25, 0.5, 1, 2, 4, 8]           # config 3
SWEEP_DEP = [0, 1, 3, 9, 27, 54, 81, 108]


def sweep_grid():
    E, D = np.meshgrid(SWEEP_EXC, SWEEP_DEP, indexing='ij')
    return E.ravel(), D.ravel()


def check_fused_reports_against_host_fit(st, monkeypatch, psf_type, steps, tol_R, min_exact):
    """The single-launch psf_report (device lmdif) against the three-launch path whose
    widths come from scipy's curve_fit, over the 64-point sweep grid: integer rescan
    ratios (array shapes) exact, resolution factors within tol_R, arrays and doses to
    1e-12; returns how many points fell back to the host fit."""
    exc, dep = sweep_grid()
    reps = st._device_reports(psf_type, exc, dep, [steps] * exc.size, [1] * exc.size, True)
    fallbacks, exact = 0, 0
    monkeypatch.setenv('LSTED_HOST_FIT', '1')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for rep, e, d in zip(reps, exc, dep):
            if rep is None:
                fallbacks += 1
                continue
            try:
                ref = st.psf_report(psf_type, e, d, steps, 1, verbose=False)
            except RuntimeError:
                # scipy gave up (maxfev): the device fit must have said so too
                raise AssertionError('device fit converged where scipy raises: %r' % ((e, d),))
            keys = ['resolution_improvement_descanned']
            if psf_type == 'line':
                keys.append('resolution_improvement_rescanned')
            for k in keys:
                assert abs(rep[k] - ref[k]) <= tol_R * abs(ref[k]), (k, e, d, rep[k], ref[k])
                exact += rep[k] == ref[k]
            for k in ('excitation_dose', 'depletion_dose', 'expected_emission'):
                assert abs(rep[k] - ref[k]) <= 1e-12 * abs(ref[k]), (k, e, d)
            assert set(rep['psfs']) == set(ref['psfs'])
            for k, v in ref['psfs'].items():
                assert rep['psfs'][k].shape == v.shape
                den = np.abs(v).max()
                assert np.abs(rep['psfs'][k] - v).max() <= 1e-12 * (den if den > 0 else 1), (k, e, d)
    monkeypatch.delenv('LSTED_HOST_FIT')
    return fallbacks, exact


def check_tune_psf_batch(st, targets):
    """tune_psf_batch == [tune_psf(**t) for t in targets], exactly."""
    one_by_one = [st.tune_psf(**t) for t in targets]
    together = st.tune_psf_batch(targets)
    assert len(together) == len(targets)
    for a, b in zip(together, one_by_one):
        assert set(a) == set(b)
        for k in b:
            if k == 'psfs':
                assert set(a[k]) == set(b[k])
                for kk in b[k]:
                    assert np.array_equal(a[k][kk], b[k][kk]), kk
            else:
                assert a[k] == b[k], k
