#!/bin/sh
# Small-shape GPU runs against the DEBUG build of the library (make debug: -DLSTED_DEBUG turns on
# the in-kernel checks of every tensor-map box, bulk-copy range / alignment and crop index --
# fft_core.cuh: LSTED_DCHECK).  compute-sanitizer is not available on the GPU pool; this is the
# once-per-round substitute.  usage (on the GPU box, from the repo root):
#     sh scripts/debug_build_tests.sh > gpurun_out/debug_build.log 2>&1
set -e
export LSTED_LIBRARY="$(pwd)/rescan_line_sted_b200/liblsted_debug.so"
test -f "$LSTED_LIBRARY" || make -C rescan_line_sted_b200/csrc debug
python scripts/sanitize_small.py
python -m pytest -q -m gpu -x tests/test_gpu_fast_path.py tests/test_gpu_deconvolver.py \
    tests/test_scan_engine.py tests/test_gpu_psf.py tests/test_orientations.py \
    "tests/test_gpu_tiled.py::test_two_tiles_against_oracle"
echo "debug build: all checks passed with $LSTED_LIBRARY"
