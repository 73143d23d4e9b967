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
        rank = 0; world = 1;
        true_object = (T*)bk.alloc(sizeof(T) * npix);
        estimate = (T*)bk.alloc(sizeof(T) * npix);
        norm = (T*)bk.alloc(sizeof(T) * npix);
        noiseless = noisy = ratio = 0;
        set_band(0, 1);
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
    void p2p_export(char*) { throw std::string("the peer-memory reduction applies to orientation sharding, not to tiles"); }
    void p2p_attach(const char*) { throw std::string("the peer-memory reduction applies to orientation sharding, not to tiles"); }
    void forget_normalization() { have_norm = false; }
    void reset_estimate() { have_estimate = false; iterations_done = 0; }
    // Sharding of a tiled object = horizontal bands ("object tiles with
    // halo"): rank r owns image rows [o0, o1) and keeps measurements / ratios on that
    // band extended by the PSF halo, where it recomputes the ratio redundantly (the
    // Poisson field is keyed by the global pixel index, so the halo copies agree with
    // the neighbour's).  The only exchange is the estimate: after every update each rank
    // broadcasts its rows of the replica (one NCCL broadcast per rank per iteration).
    void set_sharding(int rank_, int world_, int) { set_band(rank_, world_); }
    static void band_rows(int Ny, int r, int world, int* a, int* b) {
        const int base = Ny / world, extra = Ny % world;
        *a = r * base + (r < extra ? r : extra);
        *b = *a + base + (r < extra ? 1 : 0);
    }
    void set_band(int rank_, int world_) {
        if (world_ > Ny) throw std::string("more ranks than image rows");
        rank = rank_; world = world_;
        band_rows(Ny, rank, world, &o0, &o1);
        const int above = ny - 1 - sy, below = sy;   // rows H_t at row y reads: y-above .. y+below
        e0 = o0 - above < 0 ? 0 : o0 - above;
        e1 = o1 + below > Ny ? Ny : o1 + below;
        if (world == 1) { e0 = 0; e1 = Ny; }
        bpix = (size_t)(e1 - e0) * Nx;
        bk.sync();
        bk.free(noiseless); bk.free(noisy); bk.free(ratio);
        noiseless = (T*)bk.alloc(sizeof(T) * bpix * K);
        noisy = (T*)bk.alloc(sizeof(T) * bpix * K);
        ratio = (T*)bk.alloc(sizeof(T) * bpix * K);
        have_norm = false; have_estimate = false;
    }
    bool ft_error(const double*, double*) { return false; }   // tiled objects: host transform
    void info(EngineInfo* o) {
        tile.info(o);
        o->Ny = Ny; o->Nx = Nx; o->iterations_done = iterations_done;
        o->tiles_y = tiles_y; o->tiles_x = tiles_x; o->tile_out_y = out_y; o->tile_out_x = out_x;
        o->band_y0 = o0; o->band_y1 = o1;
    }

    void upload_object(const double* obj_host) { bk.upload(stage64, obj_host, sizeof(double) * npix); }
    void simulate(double total_brightness, bool rescale, unsigned long long seed) {
        double s = 1.0;
        if (rescale) s = total_brightness / bk.sum(stage64, npix, partial);
        bk.cast_in(true_object, stage64, npix, s);
        for (int ty = 0; ty < rows_tiles(e0, e1); ++ty)
            for (int tx = 0; tx < tiles_x; ++tx) {
                window(WIN_LOAD, e0, e1, ty, tx, win1, true_object, 0, 1, false);
                tile.op_H(win1, winK, 0, 0);
                WinArgs<T> a = win_args(e0, e1, ty, tx, winK, noiseless, noisy, K, true);
                a.seed = seed; a.img0 = 0;
                bk.template launch_win<WIN_SIMULATE, T>(a);
            }
    }
    void create_data(const double* obj_host, double total_brightness, bool rescale,
                     unsigned long long seed) {
        upload_object(obj_host);
        simulate(total_brightness, rescale, seed);
    }

    void ensure_norm() {
        if (have_norm) return;
        for (int ty = 0; ty < rows_tiles(o0, o1); ++ty)     // only the owned rows are ever used
            for (int tx = 0; tx < tiles_x; ++tx) {
                window(WIN_ONES, o0, o1, ty, tx, winK, 0, 0, K, false);
                tile.op_Ht_raw(winK, win2);
                window(WIN_STORE, o0, o1, ty, tx, win2, norm, 0, 1, false);
            }
        have_norm = true;
    }
    // H: big x -> big out[K] (clipped), tile by tile
    // (host-array operators: unsharded handles only; full-size K-image arrays)
    void apply_H(const T* x, T* out) {
        for (int ty = 0; ty < rows_tiles(0, Ny); ++ty)
            for (int tx = 0; tx < tiles_x; ++tx) {
                window(WIN_LOAD, 0, Ny, ty, tx, win1, const_cast<T*>(x), 0, 1, false);
                tile.op_H(win1, winK, 0, 0);
                window(WIN_STORE, 0, Ny, ty, tx, winK, out, 0, K, false);
            }
    }
    void apply_Ht_raw(const T* y, T* out) {
        for (int ty = 0; ty < rows_tiles(0, Ny); ++ty)
            for (int tx = 0; tx < tiles_x; ++tx) {
                window(WIN_LOAD, 0, Ny, ty, tx, winK, const_cast<T*>(y), 0, K, false);
                tile.op_Ht_raw(winK, win2);
                window(WIN_STORE, 0, Ny, ty, tx, win2, out, 0, 1, false);
            }
    }

    void iterate(int n) {
        for (int it = 0; it < n; ++it) {
            ensure_norm();
            if (!have_estimate) { bk.fill(estimate, npix, (T)1); have_estimate = true; }
            // ratio = measurement / H(estimate) on the owned band plus its halo
            for (int ty = 0; ty < rows_tiles(e0, e1); ++ty)
                for (int tx = 0; tx < tiles_x; ++tx) {
                    window(WIN_LOAD, e0, e1, ty, tx, win1, estimate, 0, 1, false);
                    tile.op_H(win1, winK, 0, 0);
                    window(WIN_RATIO, e0, e1, ty, tx, winK, ratio, noisy, K, true);
                }
            // estimate *= H_t(ratio) / norm on the owned band
            for (int ty = 0; ty < rows_tiles(o0, o1); ++ty)
                for (int tx = 0; tx < tiles_x; ++tx) {
                    window(WIN_LOAD, o0, o1, ty, tx, winK, ratio, 0, K, true);
                    tile.op_Ht_raw(winK, win2);
                    window(WIN_UPDATE, o0, o1, ty, tx, win2, estimate, norm, 1, false);
                }
            // every rank publishes its rows of the replicated estimate
            for (int r = 0; r < world && world > 1; ++r) {
                int a, b;
                band_rows(Ny, r, world, &a, &b);
                bk.broadcast(estimate + (size_t)a * Nx, (size_t)(b - a) * Nx, r);
            }
            ++iterations_done;
        }
    }

    // Full-size arrays are replicas; per-orientation arrays hold the band rows only
    // (host images are always full size: rows outside the band read as 0 / are ignored;
    // of H_t_normalization only the owned rows are meaningful on a sharded handle).
    void get_array(int id, int k, double* host) {
        if (id == ARR_NORMALIZATION) ensure_norm();
        if (id == ARR_NOISELESS || id == ARR_NOISY) {
            const T* src = (id == ARR_NOISY ? noisy : noiseless) + bpix * k;
            bk.fill_double(stage64, npix, 0.0);
            bk.cast_out(stage64 + (size_t)e0 * Nx, src, bpix);
        } else {
            bk.cast_out(stage64, id == ARR_TRUE_OBJECT ? true_object
                                 : id == ARR_ESTIMATE ? estimate : norm, npix);
        }
        bk.download(host, stage64, sizeof(double) * npix);
    }
    void set_array(int id, int k, const double* host) {
        bk.upload(stage64, host, sizeof(double) * npix);
        if (id == ARR_NOISELESS || id == ARR_NOISY) {
            T* dst = (id == ARR_NOISY ? noisy : noiseless) + bpix * k;
            bk.cast_in(dst, stage64 + (size_t)e0 * Nx, bpix, 1.0);
        } else {
            bk.cast_in(id == ARR_TRUE_OBJECT ? true_object : id == ARR_ESTIMATE ? estimate : norm,
                       stage64, npix, 1.0);
        }
        if (id == ARR_ESTIMATE) have_estimate = true;
        if (id == ARR_NORMALIZATION) have_norm = true;
    }
    // Host-array operators reuse `ratio` as the K-image temporary (it is rebuilt by
    // every iteration anyway).
    void H_host(const double* x, double* out) {
        if (world > 1) throw std::string("H/H_t on host arrays need an unsharded handle");
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
        if (world > 1) throw std::string("H/H_t on host arrays need an unsharded handle");
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
    size_t npix, wpix, bpix;
    int rank, world, o0, o1, e0, e1;   // owned rows [o0, o1), held rows [e0, e1)
    T *true_object, *estimate, *norm, *noiseless, *noisy, *ratio, *win1, *win2, *winK;
    double *stage64, *partial;

    int rows_tiles(int r0, int r1) const { return (r1 - r0 + out_y - 1) / out_y; }
    // Tile (ty, tx) of the sweep over image rows [r0, r1); `band` = the big arrays are
    // band arrays (rows [e0, e1)) rather than full-size replicas.
    WinArgs<T> win_args(int r0, int r1, int ty, int tx, T* tile_buf, T* big, T* big2, int nimg,
                        bool band) {
        WinArgs<T> a;
        memset(&a, 0, sizeof(a));
        a.tile = tile_buf; a.big = big; a.big2 = big2; a.nimg = nimg;
        a.W = W; a.Ny = Ny; a.Nx = Nx;
        a.y0 = r0 + ty * out_y - iy0; a.x0 = tx * out_x - ix0;
        a.iy0 = iy0; a.iy1 = iy0 + out_y; a.ix0 = ix0; a.ix1 = ix0 + out_x;
        a.by0 = band ? e0 : 0; a.brows = band ? e1 - e0 : Ny;
        a.sy0 = r0; a.sy1 = r1;
        return a;
    }
    void window(int op, int r0, int r1, int ty, int tx, T* tile_buf, T* big, T* big2, int nimg,
                bool band) {
        WinArgs<T> a = win_args(r0, r1, ty, tx, tile_buf, big, big2, nimg, band);
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
