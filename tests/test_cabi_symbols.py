"""The C-ABI library loads and exports every symbol include/lsted.h declares;
without a GPU the product fails loudly instead of falling back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from rescan_line_sted_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, 'include', 'lsted.h')) as f:
        text = re.sub(r'/\*.*?\*/', '', f.read(), flags=re.S)
    return sorted(set(re.findall(r'\b(lsted_[a-z_A-Z0-9]+)\s*\(', text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(_lib.LIBRARY_PATH):
        import __graft_entry__
        __graft_entry__.build()
    cdll = ctypes.CDLL(_lib.LIBRARY_PATH)
    for name in header_symbols():
        assert hasattr(cdll, name), name
    assert cdll.lsted_version() >= 100


def gpu_present():
    n = ctypes.c_int(0)
    lib = _lib.get()
    return lib.cdll.lsted_device_count(ctypes.byref(n)) == 0 and n.value > 0


def test_no_cpu_fallback_without_gpu(tmp_path):
    if gpu_present():
        pytest.skip('a CUDA device is present')
    from rescan_line_sted_b200 import line_sted_tools as st
    with pytest.raises(RuntimeError, match='(?i)cuda'):
        st.psf_report('point', 1, 9, 8, 1, verbose=False)
    d = st.Deconvolver([np.ones((1, 3, 3))], output_prefix=str(tmp_path) + '/x_',
                       verbose=False)
    with pytest.raises(RuntimeError, match='(?i)cuda'):
        d.create_data_from_object(np.ones((1, 8, 8)))
    with pytest.raises(RuntimeError, match='(?i)cuda'):
        d.H(np.ones((1, 8, 8)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'rescan_line_sted_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h', '.inl')):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert not re.search(r'^\s*(from|import)\s+oracle|#include[^\n]*oracle|'
                                     r'CDLL\([^\n]*oracle', text, flags=re.M), fn
