"""Worker of tests/test_distributed_gloo.py: one rank of a world_size-2 gloo
group running the orientation-sharded engine on the CPU replay backend (the
all-reduce the CUDA build gives to NCCL goes through torch.distributed here)."""
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
    from rescan_line_sted_b200 import _lib, sharded, line_sted_tools as st

    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = emul_support.emulator_library()
    _lib._library = lib   # the host mirror (psf_report_batch) also runs on the replay

    @ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_double), ctypes.c_size_t)
    def allreduce(buf, n):
        arr = np.ctypeslib.as_array(buf, shape=(n,))
        t = torch.from_numpy(arr)
        dist.all_reduce(t)          # in place: torch shares the numpy memory

    lib.cdll.emul_set_allreduce(allreduce)

    @ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_double), ctypes.c_size_t, ctypes.c_int)
    def broadcast(buf, n, root):
        arr = np.ctypeslib.as_array(buf, shape=(n,))
        dist.broadcast(torch.from_numpy(arr), src=root)

    lib.cdll.emul_set_broadcast(broadcast)

    # halo exchange of the tile-sharded engine (grouped ncclSend / ncclRecv on the GPU)
    dp, sp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_size_t)

    @ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(dp), sp,
                      ctypes.POINTER(dp), sp)
    def exchange(npeers, peers, send, nsend, recv, nrecv):
        ops, keep = [], []
        for i in range(npeers):
            if nsend[i]:
                t = torch.from_numpy(np.ctypeslib.as_array(send[i], shape=(nsend[i],)))
                ops.append(dist.P2POp(dist.isend, t, peers[i]))
            if nrecv[i]:
                t = torch.from_numpy(np.ctypeslib.as_array(recv[i], shape=(nrecv[i],)))
                ops.append(dist.P2POp(dist.irecv, t, peers[i]))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    lib.cdll.emul_set_exchange(exchange)

    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'fig2_2p0x_lr.npz'))
    out = {}
    for precision, tag in ((64, 'fp64'), (32, 'fp32')):
        d = sharded.OrientationShardedDeconvolver(g['psfs'], (128, 128), precision=precision,
                                                  lib=lib)
        assert (d.k1 - d.k0) == 4 // world
        d.create_data(g['object_u8'].astype(np.float64), 5e10, 0)
        nl = d.local_measurements(_lib.NOISELESS)
        err_nl = max(np.linalg.norm(v[0] - g['noiseless'][k]) / np.linalg.norm(g['noiseless'][k])
                     for k, v in nl.items())
        for k in range(4):
            d.set_noisy(k, g['noisy'][k])
        d.iterate(1)
        e1 = np.linalg.norm(d.estimate - g['estimate_1']) / np.linalg.norm(g['estimate_1'])
        en = np.linalg.norm(d.H_t_normalization - g['H_t_normalization']) / np.linalg.norm(
            g['H_t_normalization'])
        d.iterate(7)
        e8 = np.linalg.norm(d.estimate - g['estimate_8']) / np.linalg.norm(g['estimate_8'])
        # every rank must hold the same estimate
        t = torch.from_numpy(d.estimate.copy())
        ref = t.clone()
        dist.broadcast(ref, src=0)
        out[tag] = {'noiseless': err_nl, 'est1': e1, 'norm': en, 'est8': e8,
                    'replica_diff': float((t - ref).abs().max())}
        d.close()
    # tiled object, tiles dealt to a grid of ranks with halo exchange (config-5 style), against
    # one unsharded handle
    rng = np.random.default_rng(4)
    psfs = rng.random((3, 9, 11))
    obj = rng.random((1, 150, 170))
    single = _lib.DeconvHandle(lib, psfs, (150, 170), precision=64)
    single.create_data(obj, 1e7, 9)
    single.iterate(4)
    t = sharded.TileShardedDeconvolver(psfs, (150, 170), precision=64, lib=lib, tile_fft_len=64)
    t.create_data(obj, 1e7, 9)
    (a, b), (c, d2) = t.rows, t.cols
    noisy_same = all(np.array_equal(t.local_measurement(k)[0, a:b, c:d2],
                                    single.get(_lib.NOISY, k)[0, a:b, c:d2]) for k in range(3))
    t.iterate(4)
    est = single.get(_lib.ESTIMATE)
    out['tiles'] = {'rows': [int(a), int(b)], 'cols': [int(c), int(d2)], 'noisy_same': bool(noisy_same),
                    'est': float(np.linalg.norm(t.estimate - est) / np.linalg.norm(est))}
    t.close(), single.close()
    # sweep sharding: 6 operating points dealt round-robin, gathered in order
    exc = [0.1, 0.5, 1, 2, 4, 8]
    dep = [1, 3, 9, 27, 54, 81]
    mine = sharded.shard_items(len(exc), rank, world)
    local = st.psf_report_batch('line', [exc[i] for i in mine], [dep[i] for i in mine], 8, 1)
    local = [{k: v for k, v in r.items() if k != 'psfs'} for r in local]
    reports = sharded.gather_reports(local, len(exc))
    out['sweep_emission'] = [float(r['expected_emission']) for r in reports]
    if rank == 0:
        with open(sys.argv[1], 'w') as f:
            json.dump(out, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
