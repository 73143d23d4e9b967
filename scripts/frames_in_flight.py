#!/usr/bin/env python3
"""Throughput of config 4 (2048^2, K = 16, frame(64), fp32) with F independent frames in flight
on ONE GPU: F handles, each with its own stream; frames are issued round-robin without
synchronising in between, so the tail of one frame's kernels overlaps the next frame's.
    python scripts/frames_in_flight.py [--frames 8] [--in-flight 1 2 3]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=8)
    ap.add_argument('--in-flight', type=int, nargs='+', default=[1, 2, 3])
    ap.add_argument('--iterations', type=int, default=64)
    ap.add_argument('--interleave', type=int, default=0, help='issue iterations in slices of this many (0: whole frames)')
    args = ap.parse_args()
    from rescan_line_sted_b200 import _lib, orientations, line_sted_tools as st
    N, K = 2048, 16
    base = st.psf_report('line', verbose=False, **bench.FIG2_2P0X_LR)['psfs']['rescan_sted']
    psfs = st._stack_psfs(orientations.line_orientation_psfs(base, K, bench.EMISSION_2P0X_LR))
    obj = bench.synthetic_object(N)
    brightness = bench.total_brightness(N)
    lib = _lib.get()
    out = {}
    for F in args.in_flight:
        hs = [_lib.DeconvHandle(lib, psfs, (N, N), precision=32) for _ in range(F)]
        for h in hs:
            h.upload_object(obj)

        def issue(frames):
            if not args.interleave:
                for i in range(frames):
                    h = hs[i % F]
                    h.set_option('forget_normalization', 1)
                    h.simulate(brightness, i + 1)
                    h.iterate(args.iterations)
            else:
                for i0 in range(0, frames, F):
                    group = hs[:min(F, frames - i0)]
                    for j, h in enumerate(group):
                        h.set_option('forget_normalization', 1)
                        h.simulate(brightness, i0 + j + 1)
                    for it in range(0, args.iterations, args.interleave):
                        for h in group:
                            h.iterate(min(args.interleave, args.iterations - it))
        issue(2 * F)
        for h in hs:
            h.sync()
        t = time.perf_counter()
        issue(args.frames)
        for h in hs:
            h.sync()
        dt = time.perf_counter() - t
        out[F] = {'frames_per_s': args.frames / dt, 'ms_per_frame': 1e3 * dt / args.frames}
        for h in hs:
            h.close()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
