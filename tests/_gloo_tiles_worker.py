"""Worker of tests/test_distributed_gloo.py::test_tile_grid_world_size_4_gloo: a tiled object
whose tiles are dealt to a 2 x 2 grid of ranks (edge AND corner neighbours in the halo
exchange) on the CPU replay backend, against one unsharded handle."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import torch.distributed as dist
    import emul_support
    from rescan_line_sted_b200 import _lib, sharded

    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = emul_support.emulator_library()
    _lib._library = lib

    @ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_double), ctypes.c_size_t)
    def allreduce(buf, n):
        dist.all_reduce(torch.from_numpy(np.ctypeslib.as_array(buf, shape=(n,))))

    lib.cdll.emul_set_allreduce(allreduce)
    dp, sp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_size_t)

    @ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(dp), sp,
                      ctypes.POINTER(dp), sp)
    def exchange(npeers, peers, send, nsend, recv, nrecv):
        ops = []
        for i in range(npeers):
            if nsend[i]:
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(
                    np.ctypeslib.as_array(send[i], shape=(nsend[i],))), peers[i]))
            if nrecv[i]:
                ops.append(dist.P2POp(dist.irecv, torch.from_numpy(
                    np.ctypeslib.as_array(recv[i], shape=(nrecv[i],))), peers[i]))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    lib.cdll.emul_set_exchange(exchange)
    out = {}
    rng = np.random.default_rng(12)
    # (PSF shape, image shape, precision): odd PSF; even PSF (unequal halo above / below)
    for name, pshape, shape, precision in (('odd_fp64', (3, 9, 11), (200, 200), 64),
                                           ('even_fp64', (2, 8, 12), (190, 215), 64),
                                           ('odd_fp32', (2, 9, 11), (200, 200), 32)):
        psfs = rng.random(pshape)
        obj = rng.random((1,) + shape)
        K = pshape[0]
        single = _lib.DeconvHandle(lib, psfs, shape, precision=precision, tile_fft_len=64)
        single.create_data(obj, 1e7, 9)
        t = sharded.TileShardedDeconvolver(psfs, shape, precision=precision, lib=lib, tile_fft_len=64)
        t.create_data(obj, 1e7, 9)
        (a, b), (c, d) = t.rows, t.cols
        noisy_same = all(np.array_equal(t.local_measurement(k)[0, a:b, c:d],
                                        single.get(_lib.NOISY, k)[0, a:b, c:d]) for k in range(K))
        outside_zero = True
        for k in range(K):
            m = t.local_measurement(k).copy()
            m[0, a:b, c:d] = 0
            outside_zero = outside_zero and not m.any()
        single.iterate(3)
        t.iterate(1)
        t.iterate(2)
        est = single.get(_lib.ESTIMATE)
        mine = t.estimate
        # injected measurements (full-size host images), estimate restarted
        for k in range(K):
            img = single.get(_lib.NOISY, k) * 1.5
            single.set(_lib.NOISY, k, img)
            t.set_noisy(k, img)
        single.set_option('reset_estimate', 1)
        t.handle.set_option('reset_estimate', 1)
        single.iterate(2)
        t.iterate(2)
        est2, mine2 = single.get(_lib.ESTIMATE), t.estimate
        info = t.handle.info()
        out[name] = {'rect': [int(a), int(b), int(c), int(d)], 'noisy_same': bool(noisy_same),
                     'outside_zero': bool(outside_zero),
                     'tiles': [info.tiles_y, info.tiles_x],
                     'est': float(np.linalg.norm(mine - est) / np.linalg.norm(est)),
                     'est_injected': float(np.linalg.norm(mine2 - est2) / np.linalg.norm(est2))}
        rects = [None] * world
        dist.all_gather_object(rects, out[name]['rect'])
        out[name]['all_rects'] = rects
        t.close(), single.close()
    if rank == 0:
        with open(sys.argv[1], 'w') as f:
            json.dump(out, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
