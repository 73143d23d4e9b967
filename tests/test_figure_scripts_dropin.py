"""The reference's UNMODIFIED figure scripts run on the drop-in module (needs the
reference sources: /root/reference here, the git-ignored baseline/_ref copy on the
GPU box; plotting absorbed by a mock matplotlib).  Numbers printed / returned are
compared with the same scripts running on the reference's own line_sted_tools.
Every test runs twice: kernels replayed on the CPU (`emul`) and, marked `gpu`,
through liblsted.so on the B200 (`cuda`, fp64 and fp32)."""
import contextlib
import io
import os
import re
import runpy
import sys
from unittest import mock

import numpy as np
import pytest

import emul_support
from _reference_loader import REFERENCE_DIR, reference_available
from rescan_line_sted_b200 import _lib

pytestmark = pytest.mark.skipif(not reference_available(),
                                reason='reference sources not present')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


DEVICE = {'kind': 'emul', 'precision': 'fp64'}


@pytest.fixture(params=['emul', pytest.param('cuda-fp64', marks=pytest.mark.gpu),
                        pytest.param('cuda-fp32', marks=pytest.mark.gpu)], autouse=True)
def device(request):
    kind, _, precision = request.param.partition('-')
    DEVICE['kind'], DEVICE['precision'] = kind, precision or 'fp64'
    yield request.param
    DEVICE['kind'], DEVICE['precision'] = 'emul', 'fp64'


def loose():
    """fp32 Deconvolver: image-level tolerances of the fp32 mode (1e-5 per pass)."""
    return DEVICE['precision'] == 'fp32'


@contextlib.contextmanager
def script_environment(tmp_path, backend):
    """cwd = <tmp>/figure_generation (scripts write to ../images, ../../big_images),
    mock matplotlib, and `import line_sted_tools` resolving to `backend`."""
    work = tmp_path / backend / 'a' / 'figure_generation'
    work.mkdir(parents=True)
    (work.parent / 'images').mkdir()
    for fn in os.listdir(REFERENCE_DIR):
        if fn.endswith('.tif'):
            os.symlink(os.path.join(REFERENCE_DIR, fn), str(work / fn))
    saved_path, saved_cwd = list(sys.path), os.getcwd()
    saved_mods = {k: sys.modules.pop(k, None) for k in
                  ('line_sted_tools', 'np_tif', 'matplotlib', 'matplotlib.pyplot')}
    mpl = mock.MagicMock()
    sys.modules['matplotlib'] = mpl
    sys.modules['matplotlib.pyplot'] = mpl.pyplot
    if backend == 'b200':
        sys.path[:0] = [ROOT, REFERENCE_DIR]      # our shim first, np_tif from the reference
    else:
        sys.path[:0] = [REFERENCE_DIR]
    os.chdir(str(work))
    saved_lib = _lib._library
    if DEVICE['kind'] == 'emul':
        _lib._library = emul_support.emulator_library()
    else:
        _lib._library = None
        _lib.get()                                # the real liblsted.so (raises without it)
    os.environ['LSTED_PRECISION'] = DEVICE['precision']
    try:
        yield
    finally:
        _lib._library = saved_lib
        os.environ.pop('LSTED_PRECISION', None)
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def numbers(text):
    return [float(x) for x in re.findall(r'-?\d+\.\d+', text)]


def run_figure_1(tmp_path, backend):
    with script_environment(tmp_path, backend):
        ns = runpy.run_path(os.path.join(REFERENCE_DIR, 'line_sted_figure_1.py'),
                            run_name='figure_1')
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            for kw in (dict(psf_type='line', excitation_brightness=0.25, depletion_brightness=9,
                            steps_per_excitation_psf_width=8, pulses_per_position=2),
                       dict(psf_type='point', excitation_brightness=1, depletion_brightness=3,
                            steps_per_excitation_psf_width=6, pulses_per_position=1)):
                ns['create_figure'](**kw)
        import line_sted_tools
        return buf.getvalue(), line_sted_tools.__file__


def test_figure_1_runs_on_the_dropin(tmp_path):
    ours, where = run_figure_1(tmp_path, 'b200')
    assert where.startswith(ROOT)
    theirs, where_ref = run_figure_1(tmp_path, 'reference')
    assert where_ref.startswith(REFERENCE_DIR)
    assert 'Excitation dose:' in ours and 'Rescan STED improvement' in ours
    a, b = numbers(ours), numbers(theirs)
    assert len(a) == len(b) and len(a) > 10
    assert np.allclose(a, b, rtol=0, atol=2e-3)   # printed with 3-5 decimals
    assert [ln for ln in ours.splitlines() if not re.search(r'\d', ln)] == \
           [ln for ln in theirs.splitlines() if not re.search(r'\d', ln)]


def run_figure_2(tmp_path, backend):
    with script_environment(tmp_path, backend):
        ns = runpy.run_path(os.path.join(REFERENCE_DIR, 'line_sted_figure_2.py'),
                            run_name='figure_2')
        with contextlib.redirect_stdout(io.StringIO()):
            pair = ns['psf_comparison_pair'](
                point_resolution_improvement=1.5, line_resolution_improvement=2.68125,
                point_emissions_per_molecule=4, line_emissions_per_molecule=2.825,
                line_scan_type='descanned', line_num_orientations=3)
        import line_sted_tools as st
        import np_tif
        obj = np_tif.tif_to_array('test_object_rings.tif').astype(np.float64)
        out = {'pair': pair}
        for name in ('point_sted_psf', 'line_sted_psfs'):
            d = st.Deconvolver(pair[name], os.path.join(os.getcwd(), 'out_' + name + '_'))
            d.create_data_from_object(obj, total_brightness=5e10)
            d.record_data()
            out[name + '_noiseless'] = np.concatenate(d.noiseless_measurement)
            d.noisy_measurement = [m + 1e-9 for m in d.noiseless_measurement]  # same data both sides
            for i, save in st.logarithmic_progress(range(4), verbose=False):
                d.iterate()
                if save:
                    d.record_iteration()
            out[name + '_estimate'] = d.estimate
            out[name + '_saved'] = list(d.saved_iterations)
        return out


def test_figure_2_pipeline_runs_on_the_dropin(tmp_path):
    ours = run_figure_2(tmp_path, 'b200')
    theirs = run_figure_2(tmp_path, 'reference')
    for which in ('point', 'line'):
        for k in ('excitation_dose', 'depletion_dose', 'expected_emission',
                  'excitation_brightness', 'depletion_brightness', 'pulses_per_position'):
            assert ours['pair'][which][k] == pytest.approx(theirs['pair'][which][k], rel=1e-5), k
    rel = lambda a, b: np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))
    for a, b in zip(ours['pair']['line_sted_psfs'], theirs['pair']['line_sted_psfs']):
        assert rel(a, b) < 1e-4      # operating point found by Brent to ~1e-6
    for name in ('point_sted_psf', 'line_sted_psfs'):
        assert rel(ours[name + '_noiseless'], theirs[name + '_noiseless']) < 1e-4
        assert rel(ours[name + '_estimate'], theirs[name + '_estimate']) < 1e-3
        assert ours[name + '_saved'] == theirs[name + '_saved']


def test_figure_a1_sweep_runs_on_the_dropin(tmp_path):
    results = {}
    for backend in ('b200', 'reference'):
        with script_environment(tmp_path, backend):
            ns = runpy.run_path(os.path.join(REFERENCE_DIR, 'line_sted_figure_a1.py'),
                                run_name='figure_a1')
            plt = sys.modules['matplotlib.pyplot']
            ns['r_vs_depletion_brightness']()
            curves = [c.args[1] for c in plt.plot.call_args_list]
            results[backend] = np.array(curves, dtype=float)
    assert results['b200'].shape == (2, 75)
    assert np.allclose(results['b200'], results['reference'], rtol=1e-6)
