"""Multi-GPU sharding of the hot path along its independent axes
(SURVEY.md section 8e), one process per GPU with torch.distributed as the
plumbing:

* `OrientationShardedDeconvolver` -- the K line orientations of ONE frame
  are split over the ranks; every rank keeps a replica of the estimate and
  one NCCL all-reduce per RL iteration (inside liblsted, on the handle's
  stream, over NVLink) sums the Fourier-domain partial H_t.
* `shard_items` / `gather_reports` -- independent units (frames, sweep
  points of `psf_report_batch`) dealt round-robin, results gathered once.
"""
import ctypes
import os

import numpy as np

from . import _lib


def orientation_slice(K, rank, world):
    """Contiguous block of orientation indices owned by `rank`."""
    if world > K:
        raise ValueError('more ranks (%d) than orientations (%d)' % (world, K))
    base, extra = divmod(K, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_items(n_items, rank, world):
    """Indices of the independent work items (frames, sweep points) of `rank`."""
    return list(range(rank, n_items, world))


def gather_reports(local, n_items, group=None):
    """Inverse of `shard_items`: every rank gets the full, ordered list."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    out = [None] * n_items
    for r, part in enumerate(parts):
        for i, item in zip(range(r, n_items, world), part):
            out[i] = item
    return out


class OrientationShardedDeconvolver:
    """Forward model + multi-view Richardson-Lucy with the orientations of one
    frame sharded over the ranks of a torch.distributed process group.

    Every rank passes the FULL list of PSFs and the same object / seed; each
    keeps only its slice on its GPU.  Results (`estimate`) are identical on
    all ranks and equal to the single-GPU result up to summation order.
    """

    def __init__(self, psfs, image_shape, precision=32, device=0, group=None,
                 lib=None):
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        psfs = np.ascontiguousarray(psfs, dtype=np.float64)
        self.K_total = psfs.shape[0]
        self.k0, self.k1 = orientation_slice(self.K_total, self.rank, self.world)
        lib = lib or _lib.get()
        self.handle = _lib.DeconvHandle(lib, psfs[self.k0:self.k1], image_shape,
                                        precision=precision, device=device)
        unique_id = [None]
        if self.world > 1:
            if self.rank == 0:
                unique_id[0] = _lib.nccl_unique_id(lib)
            dist.broadcast_object_list(unique_id, src=0, group=group)
        self.handle.shard(self.rank, self.world, self.k0, unique_id[0])
        info = self.handle.info()
        # fp32: the per-iteration sum over ranks goes THROUGH THE NVSWITCH (NVLS): the spectrum of
        # the partial sums is bound to one CUDA multicast object; multimem.ld_reduce adds the
        # replicas in the switch, multimem.st copies the result into all of them.  Measured on
        # B200 (2048^2, K = 16): 8 GPUs 85.2 vs 72.1 frames/s with the peer-memory kernel below,
        # 2 GPUs 44.9 vs 46.8 (a rank's own replica also travels through the switch), so the
        # default is NVLS from 4 ranks up; LSTED_NVLS=1 / 0 forces / forbids it.  Boxes without
        # multicast support fall through to the paths below.
        self.nvls = False
        want_nvls = os.environ.get('LSTED_NVLS', 'auto')
        if (self.world > 1 and precision == 32 and info.tiles_y * info.tiles_x == 1
                and (want_nvls == '1' or (want_nvls != '0' and self.world >= 4))):
            self.nvls = self._attach_nvls(lib, device, group)
        # Fast path (2160-point transforms): the per-iteration sum over ranks runs inside the
        # column kernel over NVLink peer memory; the CUDA-IPC handles travel through the
        # process group.  LSTED_P2P=0 keeps the NCCL all-reduce.
        self.p2p = False
        if (self.world > 1 and not self.nvls and os.environ.get('LSTED_P2P', '1') != '0'
                and info.Ly == 2160 and info.tiles_y * info.tiles_x == 1):
            mine = self.handle.p2p_export()
            parts = [None] * self.world
            dist.all_gather_object(parts, mine, group=group)
            self.handle.p2p_attach(b''.join(parts), self.world)
            dist.barrier(group=group)
            self.p2p = True

    def _attach_nvls(self, lib, device, group):
        """Bind this rank's partial-sum spectrum to a multicast object shared by the group.
        The object is created on rank 0; its POSIX file descriptor reaches the other
        processes as SCM_RIGHTS ancillary data over a Unix socket (the one way to move a
        descriptor between unrelated processes).  Every stage is agreed on by all ranks: a
        failure anywhere before the bind makes every rank fall back to the other reductions."""
        import socket
        import torch.distributed as dist

        def everyone(ok):
            flags = [None] * self.world
            dist.all_gather_object(flags, bool(ok), group=group)
            return all(flags)

        ok = ctypes.c_int(0)
        try:
            lib.call('lsted_deconv_nvls_supported', int(device), ctypes.byref(ok))
        except RuntimeError:
            ok = ctypes.c_int(0)
        if not everyone(ok.value):
            return False
        path, server, fd, good = [None], None, -1, True
        if self.rank == 0:
            try:
                fd = self.handle.nvls_create(self.world)
                path[0] = '\0lsted_nvls_%d_%d' % (os.getpid(), id(self))    # abstract socket name
                server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                server.bind(path[0])
                server.listen(self.world)
                server.settimeout(60)
            except (RuntimeError, OSError):
                path[0], good = None, False
        dist.broadcast_object_list(path, src=0, group=group)
        if path[0] is None:
            return False
        try:
            if self.rank == 0:
                for _ in range(self.world - 1):
                    conn, _ = server.accept()
                    socket.send_fds(conn, [b'mc'], [fd])
                    conn.settimeout(60)
                    conn.recv(2)                     # the peer has imported the object (or given up)
                    conn.close()
            else:
                c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                c.settimeout(60)
                c.connect(path[0])
                _, fds, _, _ = socket.recv_fds(c, 16, 1)
                try:
                    self.handle.nvls_import(self.world, fds[0])
                except RuntimeError:
                    good = False
                c.send(b'ok')
                c.close()
                os.close(fds[0])
        except OSError:
            good = False
        finally:
            if server is not None:
                server.close()
            if fd >= 0:
                os.close(fd)
        if not everyone(good):
            return False
        try:
            self.handle.nvls_add_device()
        except RuntimeError:
            good = False
        if not everyone(good):                   # (also the barrier: every device is in the team)
            return False
        try:
            self.handle.nvls_bind()
        except RuntimeError:
            good = False
        if not everyone(good):                   # (also the barrier: every replica is bound)
            raise RuntimeError('NVLS: binding the multicast object failed on a rank; set LSTED_NVLS=0')
        return True

    def create_data(self, obj, total_brightness, seed):
        """Each rank simulates only its orientations; the Poisson streams are
        keyed by the global orientation index, so the noise field does not
        depend on the number of GPUs."""
        self.handle.create_data(obj, total_brightness, seed)

    def set_noisy(self, k_global, image):
        if self.k0 <= k_global < self.k1:
            self.handle.set(_lib.NOISY, k_global - self.k0, image)

    def local_measurements(self, which=_lib.NOISY):
        return {k: self.handle.get(which, k - self.k0) for k in range(self.k0, self.k1)}

    def iterate(self, n=1):
        self.handle.iterate(n)

    @property
    def estimate(self):
        return self.handle.get(_lib.ESTIMATE)

    @property
    def H_t_normalization(self):
        return self.handle.get(_lib.NORMALIZATION)

    def close(self):
        self.handle.close()


class TileShardedDeconvolver:
    """A large object processed in overlap-save tiles, the tiles dealt to a 2-D grid of
    ranks (BASELINE config 5: 4 x 4 tiles of 8192^2 on a 2 x 4 grid of 8 GPUs).  Every rank
    passes the same PSFs, object and seed; measurements and ratios exist only on the rank's
    rectangle `rows` x `cols`; per RL iteration the PSF-halo ring of the estimate and of the K
    ratio images is exchanged with the neighbouring ranks (grouped ncclSend / ncclRecv).
    `estimate` assembles the full image from the owners: call it on every rank."""

    def __init__(self, psfs, image_shape, precision=64, device=0, group=None,
                 lib=None, tile_fft_len=2160):
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        lib = lib or _lib.get()
        self.handle = _lib.DeconvHandle(lib, psfs, image_shape, precision=precision,
                                        device=device, tile_fft_len=tile_fft_len)
        unique_id = [None]
        if self.world > 1:
            if self.rank == 0:
                unique_id[0] = _lib.nccl_unique_id(lib)
            dist.broadcast_object_list(unique_id, src=0, group=group)
        self.handle.shard(self.rank, self.world, 0, unique_id[0])
        info = self.handle.info()
        self.rows = (info.band_y0, info.band_y1)
        self.cols = (info.band_x0, info.band_x1)

    def create_data(self, obj, total_brightness, seed):
        self.handle.create_data(obj, total_brightness, seed)

    def set_noisy(self, k, image):
        """Full-size image; the rank keeps its rectangle."""
        self.handle.set(_lib.NOISY, k, image)

    def iterate(self, n=1):
        self.handle.iterate(n)

    @property
    def estimate(self):
        return self.handle.get(_lib.ESTIMATE)

    def local_measurement(self, k, which=_lib.NOISY):
        """The rank's rectangle of the image; zeros elsewhere."""
        return self.handle.get(which, k)

    def close(self):
        self.handle.close()
