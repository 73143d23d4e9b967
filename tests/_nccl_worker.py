"""Worker of tests/test_gpu_multi.py: one rank (one GPU) of an NCCL group."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from rescan_line_sted_b200 import _lib, sharded
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'fig2_2p0x_lr.npz'))
    out = {}
    for precision, tag in ((64, 'fp64'), (32, 'fp32')):
        d = sharded.OrientationShardedDeconvolver(g['psfs'], (128, 128), precision=precision,
                                                  device=local)
        d.create_data(g['object_u8'].astype(np.float64), 5e10, 0)
        for k in range(4):
            d.set_noisy(k, g['noisy'][k])
        d.iterate(1)
        e1 = np.linalg.norm(d.estimate - g['estimate_1']) / np.linalg.norm(g['estimate_1'])
        d.iterate(7)
        est = d.estimate
        e8 = np.linalg.norm(est - g['estimate_8']) / np.linalg.norm(g['estimate_8'])
        t = torch.from_numpy(est.copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        out[tag] = {'est1': e1, 'est8': e8, 'replica_diff': float((t - ref).abs().max()),
                    'nvls': bool(d.nvls)}
        # in-kernel Poisson field must not depend on the GPU count
        d.create_data(g['object_u8'].astype(np.float64), 5e10, 123)
        noisy = d.local_measurements(_lib.NOISY)
        single = _lib.DeconvHandle(_lib.get(), g['psfs'], (128, 128), precision=precision,
                                   device=local)
        single.create_data(g['object_u8'].astype(np.float64), 5e10, 123)
        same = all(np.array_equal(v, single.get(_lib.NOISY, k)) for k, v in noisy.items())
        out[tag]['noise_independent_of_world'] = bool(same)
        single.close()
        d.close()
    # 2160-point fast path: the sum over ranks runs inside the column kernel over NVLink peer
    # memory (CUDA IPC); against one GPU and against the NCCL all-reduce variant
    rng = np.random.default_rng(11)
    psfs = rng.random((6, 9, 11))
    obj = rng.random((1, 2100, 2100)) + 0.1
    for precision, tag, tol in ((32, 'p2p_fp32', 1e-5), (64, 'p2p_fp64', 1e-12)):
        if precision == 64:   # point-symmetric PSFs: centred real OTFs on one GPU, the fused
            psfs = 0.5 * (psfs + psfs[:, ::-1, ::-1])   # reduction reads the centred complex ones
        single = _lib.DeconvHandle(_lib.get(), psfs, (2100, 2100), precision=precision, device=local)
        single.create_data(obj, 1e10, 3)
        single.iterate(3)
        want = single.get(_lib.ESTIMATE)
        res = {}
        # reductions: 'nvls' multicast through the switch (fp32), '1' peer-memory kernel, '0' NCCL
        for mode in (('nvls',) if precision == 32 else ()) + ('1', '0'):
            os.environ['LSTED_NVLS'] = '1' if mode == 'nvls' else '0'
            os.environ['LSTED_P2P'] = '0' if mode == '0' else '1'
            d = sharded.OrientationShardedDeconvolver(psfs, (2100, 2100), precision=precision,
                                                      device=local)
            if mode == 'nvls' and not d.nvls:      # no multicast support on this box
                res['nvls_unavailable'] = True
                d.close()
                continue
            assert d.nvls == (mode == 'nvls') and d.p2p == (mode == '1')
            d.create_data(obj, 1e10, 3)
            d.iterate(3)
            est = d.estimate
            res[mode] = float(np.linalg.norm(est - want) / np.linalg.norm(want))
            t = torch.from_numpy(est.copy()).cuda()
            ref = t.clone()
            dist.broadcast(ref, src=0)
            res[mode + '_replica_diff'] = float((t - ref).abs().max() / ref.abs().max())
            d.close()
        os.environ.pop('LSTED_P2P', None)
        os.environ.pop('LSTED_NVLS', None)
        out[tag] = dict(res, tol=tol)
        single.close()
    # tiled object, tiles dealt to the GPUs with halo exchange, 2160-point tiles (fast path), fp64
    rng = np.random.default_rng(4)
    psfs = rng.random((2, 9, 11))
    obj = rng.random((1, 2300, 2200))
    single = _lib.DeconvHandle(_lib.get(), psfs, (2300, 2200), precision=64, device=local,
                               tile_fft_len=2160)
    single.create_data(obj, 1e9, 9)
    single.iterate(2)
    t = sharded.TileShardedDeconvolver(psfs, (2300, 2200), precision=64, device=local)
    t.create_data(obj, 1e9, 9)
    (a, b), (c, d2) = t.rows, t.cols
    same = all(np.array_equal(t.local_measurement(k)[0, a:b, c:d2],
                              single.get(_lib.NOISY, k)[0, a:b, c:d2]) for k in range(2))
    t.iterate(2)
    est = single.get(_lib.ESTIMATE)
    out['tiles'] = {'noisy_same': bool(same), 'rect': [int(a), int(b), int(c), int(d2)],
                    'est': float(np.linalg.norm(t.estimate - est) / np.linalg.norm(est))}
    t.close(), single.close()
    if rank == 0:
        with open(sys.argv[1], 'w') as f:
            json.dump(out, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
