// Overlap-save tiling of the convolution engine for objects whose padded
// transform does not fit a CTA's shared memory (BASELINE config 5: 8192^2).
//
// The object is cut into tiles; each tile is convolved through a W x W window
// (W = the tile FFT length, 2160 by default so the compile-time plan applies)
// that carries the (n-1)-pixel halo, zero-filled outside the object exactly like
// the reference's linear 'same' convolution; only the alias-free interior
// W - (n-1) of a window is written back.  H and H_t therefore equal the
// untiled operators (tests compare them), and one RL iteration becomes two
// sweeps over the tiles: ratio = measurement / H(estimate), then
// estimate *= H_t(ratio) / norm (the ratio is materialised because its halo
// comes from the neighbouring tiles).
//
// Mirrors Deconvolver (figure_generation/line_sted_tools.py:478-594) like
// DeconvEngine does; full-size arrays live in HBM (3 x K x Ny x Nx elements
// for measurements and ratios: 52 GB at 8192^2, K = 32, fp64).
#pragma once
#include "engine.h"
#include "ew_bodies.cuh"

namespace lsted {

template <typename T, class BK> class TiledEngine : public EngineBase {
  public:
    TiledEngine(BK& backend, int K_, int ny_, int nx_, int Ny_, int Nx_, int tile_L)
        : K(K_), ny(ny_), nx(nx_), Ny(Ny_), Nx(Nx_), W(tile_L), iterations_done(0),
          have_norm(false), have_estimate(false), bk(backend),
          tile(backend, K_, ny_, nx_, tile_L, tile_L, tile_L) {
        if (W - (ny - 1) < 1 || W - (nx - 1) < 1) throw std::string("tile FFT length shorter than the PSF");
        sy = (ny - 1) / 2; sx = (nx - 1) / 2;
        iy0 = ny - 1 - sy; ix0 = nx - 1 - sx;     // first alias-free output of a window
        out_y = W - (ny - 1); out_x = W - (nx - 1);
        tiles_y = (Ny + out_y - 1) / out_y; tiles_x = (Nx + out_x - 1) / out_x;
        npix = (size_t)Ny * Nx;
        wpix = (size_t)W * W;
        true_object = (T*)bk.alloc(sizeof(T) * npix);
        estimate = (T*)bk.alloc(sizeof(T) * npix);
        norm = (T*)bk.alloc(sizeof(T) * npix);
        noiseless = (T*)bk.alloc(sizeof(T) * npix * K);
        noisy = (T*)bk.alloc(sizeof(T) * npix * K);
        ratio = (T*)bk.alloc(sizeof(T) * npix * K);
        win1 = (T*)bk.alloc(sizeof(T) * wpix);
        win2 = (T*)bk.alloc(sizeof(T) * wpix);
        winK = (T*)bk.alloc(sizeof(T) * wpix * K);
        stage64 = (double*)bk.alloc(sizeof(double) * npix);
        partial = (double*)bk.alloc(sizeof(double) * 1024);
    }
    ~TiledEngine() {
        void* all[] = {true_object, estimate, norm, noiseless, noisy, ratio, win1, win2, winK,
                       stage64, partial};
        for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i) bk.free(all[i]);
    }

    void set_psfs(const double* psfs_host) { tile.set_psfs(psfs_host); have_norm = false; }
    void set_exact_clip(bool on) { tile.set_exact_clip(on); have_norm = false; }
    void forget_normalization() { have_norm = false; }
    void set_sharding(int, int world, int) {
        if (world > 1) throw std::string("orientation sharding of a tiled object is not implemented");
    }
    void info(EngineInfo* o) {
        tile.info(o);
        o->Ny = Ny; o->Nx = Nx; o->iterations_done = iterations_done;
        o->tiles_y = tiles_y; o->tiles_x = tiles_x; o->tile_out_y = out_y; o->tile_out_x = out_x;
    }

    void upload_object(const double* obj_host) { bk.upload(stage64, obj_host, sizeof(double) * npix); }
    void simulate(double total_brightness, bool rescale, unsigned long long seed) {
        double s = 1.0;
        if (rescale) s = total_brightness / bk.sum(stage64, npix, partial);
        bk.cast_in(true_object, stage64, npix, s);
        for (int ty = 0; ty < tiles_y; ++ty)
            for (int tx = 0; tx < tiles_x; ++tx) {
                window(WIN_LOAD, ty, tx, win1, true_object, 0, 1);
                tile.op_H(win1, winK, 0, 0);
                WinArgs<T> a = win_args(ty, tx, winK, noiseless, noisy, K);
                a.seed = seed; a.img0 = 0;
                bk.template launch_win<WIN_SIMULATE, T>(a);
            }
        iterations_done = 0;
        have_estimate = false;
    }
    void create_data(const double* obj_host, double total_brightness, bool rescale,
                     unsigned long long seed) {
        upload_object(obj_host);
        simulate(total_brightness, rescale, seed);
    }

    void ensure_norm() {
        if (have_norm) return;
        for (int ty = 0; ty < tiles_y; ++ty)
            for (int tx = 0; tx < tiles_x; ++tx) {
                window(WIN_ONES, ty, tx, winK, 0, 0, K);
                tile.op_Ht_raw(winK, win2);
                window(WIN_STORE, ty, tx, win2, norm, 0, 1);
            }
        have_norm = true;
    }
    // H: big x -> big out[K] (clipped), tile by tile
    void apply_H(const T* x, T* out) {
        for (int ty = 0; ty < tiles_y; ++ty)
            for (int tx = 0; tx < tiles_x; ++tx) {
                window(WIN_LOAD, ty, tx, win1, const_cast<T*>(x), 0, 1);
                tile.op_H(win1, winK, 0, 0);
                window(WIN_STORE, ty, tx, winK, out, 0, K);
            }
    }
    void apply_Ht_raw(const T* y, T* out) {
        for (int ty = 0; ty < tiles_y; ++ty)
            for (int tx = 0; tx < tiles_x; ++tx) {
                window(WIN_LOAD, ty, tx, winK, const_cast<T*>(y), 0, K);
                tile.op_Ht_raw(winK, win2);
                window(WIN_STORE, ty, tx, win2, out, 0, 1);
            }
    }

    void iterate(int n) {
        for (int it = 0; it < n; ++it) {
            ensure_norm();
            if (!have_estimate) { bk.fill(estimate, npix, (T)1); have_estimate = true; }
            for (int ty = 0; ty < tiles_y; ++ty)       // ratio = measurement / H(estimate)
                for (int tx = 0; tx < tiles_x; ++tx) {
                    window(WIN_LOAD, ty, tx, win1, estimate, 0, 1);
                    tile.op_H(win1, winK, 0, 0);
                    window(WIN_RATIO, ty, tx, winK, ratio, noisy, K);
                }
            for (int ty = 0; ty < tiles_y; ++ty)       // estimate *= H_t(ratio) / norm
                for (int tx = 0; tx < tiles_x; ++tx) {
                    window(WIN_LOAD, ty, tx, winK, ratio, 0, K);
                    tile.op_Ht_raw(winK, win2);
                    window(WIN_UPDATE, ty, tx, win2, estimate, norm, 1);
                }
            ++iterations_done;
        }
    }

    T* array(int id, int k) {
        switch (id) {
            case ARR_TRUE_OBJECT: return true_object;
            case ARR_NOISELESS: return noiseless + npix * k;
            case ARR_NOISY: return noisy + npix * k;
            case ARR_ESTIMATE: return estimate;
            case ARR_NORMALIZATION: return norm;
        }
        return 0;
    }
    void get_array(int id, int k, double* host) {
        if (id == ARR_NORMALIZATION) ensure_norm();
        bk.cast_out(stage64, array(id, k), npix);
        bk.download(host, stage64, sizeof(double) * npix);
    }
    void set_array(int id, int k, const double* host) {
        bk.upload(stage64, host, sizeof(double) * npix);
        bk.cast_in(array(id, k), stage64, npix, 1.0);
        if (id == ARR_ESTIMATE) have_estimate = true;
        if (id == ARR_NORMALIZATION) have_norm = true;
    }
    // Host-array operators reuse `ratio` as the K-image temporary (it is rebuilt by
    // every iteration anyway).
    void H_host(const double* x, double* out) {
        T* xin = (T*)bk.alloc(sizeof(T) * npix);
        bk.upload(stage64, x, sizeof(double) * npix);
        bk.cast_in(xin, stage64, npix, 1.0);
        apply_H(xin, ratio);
        for (int k = 0; k < K; ++k) {
            bk.cast_out(stage64, ratio + npix * k, npix);
            bk.download(out + npix * k, stage64, sizeof(double) * npix);
        }
        bk.free(xin);
    }
    void Ht_host(const double* y, double* out, bool normalize) {
        if (normalize) ensure_norm();
        T* res = (T*)bk.alloc(sizeof(T) * npix);
        for (int k = 0; k < K; ++k) {
            bk.upload(stage64, y + npix * k, sizeof(double) * npix);
            bk.cast_in(ratio + npix * k, stage64, npix, 1.0);
        }
        apply_Ht_raw(ratio, res);
        if (normalize) bk.divide(res, norm, npix);
        bk.cast_out(stage64, res, npix);
        bk.download(out, stage64, sizeof(double) * npix);
        bk.free(res);
    }

  private:
    int K, ny, nx, Ny, Nx, W;
    int sy, sx, iy0, ix0, out_y, out_x, tiles_y, tiles_x;
    int iterations_done;
    bool have_norm, have_estimate;
    BK& bk;
    DeconvEngine<T, BK> tile;
    size_t npix, wpix;
    T *true_object, *estimate, *norm, *noiseless, *noisy, *ratio, *win1, *win2, *winK;
    double *stage64, *partial;

    WinArgs<T> win_args(int ty, int tx, T* tile_buf, T* big, T* big2, int nimg) {
        WinArgs<T> a;
        memset(&a, 0, sizeof(a));
        a.tile = tile_buf; a.big = big; a.big2 = big2; a.nimg = nimg;
        a.W = W; a.Ny = Ny; a.Nx = Nx;
        a.y0 = ty * out_y - iy0; a.x0 = tx * out_x - ix0;
        a.iy0 = iy0; a.iy1 = iy0 + out_y; a.ix0 = ix0; a.ix1 = ix0 + out_x;
        return a;
    }
    void window(int op, int ty, int tx, T* tile_buf, T* big, T* big2, int nimg) {
        WinArgs<T> a = win_args(ty, tx, tile_buf, big, big2, nimg);
        switch (op) {
            case WIN_LOAD: bk.template launch_win<WIN_LOAD, T>(a); break;
            case WIN_ONES: bk.template launch_win<WIN_ONES, T>(a); break;
            case WIN_STORE: bk.template launch_win<WIN_STORE, T>(a); break;
            case WIN_RATIO: bk.template launch_win<WIN_RATIO, T>(a); break;
            default: bk.template launch_win<WIN_UPDATE, T>(a); break;
        }
    }
};

}  // namespace lsted
