"""Import the UNMODIFIED reference module from /root/reference (build
container) or from its git-ignored copy baseline/_ref/ that __graft_entry__.build()
makes and gpurun ships to the GPU box.  Used by the golden-vector generator
and by the CPU tests that pin the oracle.  matplotlib is not installed in the
image and the reference imports (but never uses) it, so a stub is injected."""
import importlib.util
import os
import sys
import types
import warnings

_CANDIDATES = ('/root/reference/figure_generation',
               os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                            'baseline', '_ref'))
REFERENCE_DIR = os.environ.get('LSTED_REFERENCE_DIR') or next(
    (d for d in _CANDIDATES if os.path.isfile(os.path.join(d, 'line_sted_tools.py'))),
    _CANDIDATES[0])


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'line_sted_tools.py'))


def load_reference():
    if 'matplotlib' not in sys.modules:
        mpl = types.ModuleType('matplotlib')
        mpl.pyplot = types.ModuleType('matplotlib.pyplot')
        sys.modules['matplotlib'] = mpl
        sys.modules['matplotlib.pyplot'] = mpl.pyplot
    if 'np_tif' not in sys.modules:
        spec = importlib.util.spec_from_file_location(
            'np_tif', os.path.join(REFERENCE_DIR, 'np_tif.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules['np_tif'] = mod
    spec = importlib.util.spec_from_file_location(
        '_reference_line_sted_tools',
        os.path.join(REFERENCE_DIR, 'line_sted_tools.py'))
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        spec.loader.exec_module(mod)
    return mod
