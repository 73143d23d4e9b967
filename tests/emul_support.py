"""Test-only helper: point the ctypes layer at the CPU replay of the kernel
bodies (tests/host_emul) so the host-side mirror can be exercised without a
GPU.  Never used by the product; GPU tests use the real liblsted.so."""
import os
import subprocess

from rescan_line_sted_b200 import _lib

EMUL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'host_emul')
EMUL_LIB = os.path.join(EMUL_DIR, 'liblsted_emul.so')


def build_emulator():
    src = os.path.join(EMUL_DIR, 'emul.cpp')
    csrc = os.path.join(os.path.dirname(EMUL_DIR), os.pardir,
                        'rescan_line_sted_b200', 'csrc')
    newest = max([os.path.getmtime(src)] +
                 [os.path.getmtime(os.path.join(csrc, f))
                  for f in os.listdir(csrc)
                  if f.endswith(('.cuh', '.h', '.inl'))])
    if not os.path.isfile(EMUL_LIB) or os.path.getmtime(EMUL_LIB) < newest:
        subprocess.check_call(['sh', os.path.join(EMUL_DIR, 'build.sh')])
    return EMUL_LIB


def emulator_library():
    sigs = dict(_lib._DECONV_SIGNATURES)
    for name in ('lsted_psf_illumination', 'lsted_psf_rescan', 'lsted_psf_rotate',
                 'lsted_psf_report_batch', 'lsted_gauss_fit'):
        sigs[name] = _lib._CORE_SIGNATURES[name]
    sigs.update(_lib._SCAN_SIGNATURES)
    return _lib.Library(build_emulator(), sigs)
