// extern "C" Deconvolver entry points of include/lsted.h, written once over
// the backend type LSTED_BACKEND (CudaBackend in lsted_api.cu; the CPU
// replay backend in tests/host_emul/emul.cpp, test infrastructure only).
// The including file provides: LSTED_BACKEND, set_error(), g_error.

struct lsted_deconv {
    int precision;
    int device;
    LSTED_BACKEND* bk;
    lsted::EngineBase* e;
};

#define LSTED_TRY try {
#define LSTED_CATCH                                                        \
    }                                                                      \
    catch (const lsted::ApiError& e) { return set_error(e.code, e.msg); }  \
    catch (const std::string& s) { return set_error(LSTED_ERR_ARG, s); }   \
    catch (const std::bad_alloc&) { return set_error(LSTED_ERR_ARG, "host allocation failed"); } \
    catch (...) { return set_error(LSTED_ERR_ARG, "unexpected exception"); }
#define LSTED_ENGINE(h, call)  \
    do {                       \
        (h)->bk->activate();   \
        (h)->e->call;          \
    } while (0)

enum { kDefaultTileFftLen = 2160 };

template <typename T>
static lsted::EngineBase* make_engine(LSTED_BACKEND& bk, int K, int ny, int nx, int Ny, int Nx,
                                      int tile_fft_len) {
    if (tile_fft_len > 0)
        return new lsted::TiledEngine<T, LSTED_BACKEND>(bk, K, ny, nx, Ny, Nx, tile_fft_len);
    try {
        return new lsted::DeconvEngine<T, LSTED_BACKEND>(bk, K, ny, nx, Ny, Nx);
    } catch (const std::string& why) {
        // too large for one shared-memory transform: overlap-save tiles
        if (why.find("tile the object") == std::string::npos) throw;
        return new lsted::TiledEngine<T, LSTED_BACKEND>(bk, K, ny, nx, Ny, Nx, kDefaultTileFftLen);
    }
}

extern "C" int lsted_deconv_create_tiled(lsted_deconv** out, int device, const double* psfs, int K,
                                         int ny, int nx, int Ny, int Nx, int precision,
                                         int tile_fft_len) {
    if (!out || !psfs) return set_error(LSTED_ERR_ARG, "null pointer");
    if (precision != 32 && precision != 64) return set_error(LSTED_ERR_ARG, "precision must be 32 or 64");
    if (K < 1) return set_error(LSTED_ERR_ARG, "need at least one PSF");
    if (tile_fft_len < 0) return set_error(LSTED_ERR_ARG, "negative tile FFT length");
    lsted_deconv* h = new lsted_deconv();
    h->precision = precision; h->device = device; h->bk = 0; h->e = 0;
    LSTED_TRY
    h->bk = new LSTED_BACKEND(device);
    h->bk->activate();
    h->e = precision == 32 ? make_engine<float>(*h->bk, K, ny, nx, Ny, Nx, tile_fft_len)
                           : make_engine<double>(*h->bk, K, ny, nx, Ny, Nx, tile_fft_len);
    h->e->set_psfs(psfs);
    *out = h;
    return LSTED_OK;
    }
    catch (const lsted::ApiError& e) { lsted_deconv_destroy(h); return set_error(e.code, e.msg); }
    catch (const std::string& s) { lsted_deconv_destroy(h); return set_error(LSTED_ERR_ARG, s); }
    catch (const std::bad_alloc&) { lsted_deconv_destroy(h); return set_error(LSTED_ERR_ARG, "host allocation failed"); }
    catch (...) { lsted_deconv_destroy(h); return set_error(LSTED_ERR_ARG, "unexpected exception in lsted_deconv_create"); }
}

extern "C" int lsted_deconv_create(lsted_deconv** out, int device, const double* psfs, int K, int ny,
                                   int nx, int Ny, int Nx, int precision) {
    return lsted_deconv_create_tiled(out, device, psfs, K, ny, nx, Ny, Nx, precision, 0);
}

extern "C" int lsted_deconv_destroy(lsted_deconv* h) {
    if (!h) return LSTED_OK;
    try {
        if (h->bk) h->bk->activate();
        delete h->e;
        delete h->bk;
    } catch (...) {}
    delete h;
    return LSTED_OK;
}

extern "C" int lsted_deconv_info(lsted_deconv* h, lsted_deconv_info_t* info) {
    if (!h || !info) return set_error(LSTED_ERR_ARG, "null pointer");
    memset(info, 0, sizeof(*info));
    LSTED_TRY
    h->bk->activate();
    lsted::EngineInfo ei;
    h->e->info(&ei);
    const int cb = h->precision == 32 ? 8 : 16;
    info->K = ei.K; info->ny = ei.ny; info->nx = ei.nx;
    info->Ny = ei.Ny; info->Nx = ei.Nx; info->Ly = ei.Ly; info->Lx = ei.Lx;
    info->cols_per_cta = ei.C; info->row_pairs_per_cta = ei.PR;
    info->precision = h->precision;
    info->iterations_done = ei.iterations_done;
    info->device = h->device;
    info->row_smem_bytes = ei.row_smem;
    info->col_smem_bytes = ei.col_smem;
    info->device_bytes = h->bk->bytes_allocated();
    const double A = (double)(cb / 2) * ei.Ny * ei.Nx;
    info->bytes_forward = (2.0 * ei.K + 1.0) * A;
    info->bytes_normalization = A;
    info->bytes_iteration = (ei.K + 4.0) * A;
    info->tiles_y = ei.tiles_y; info->tiles_x = ei.tiles_x;
    info->tile_out_y = ei.tile_out_y; info->tile_out_x = ei.tile_out_x;
    info->band_y0 = ei.band_y0; info->band_y1 = ei.band_y1;
    info->band_x0 = ei.band_x0; info->band_x1 = ei.band_x1;
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_set_option(lsted_deconv* h, const char* name, double value) {
    if (!h || !name) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    h->bk->activate();
    // switches that select kernels or their arguments invalidate the captured iteration
    static const char* const kernel_switches[] = {"prefetch", "fast_path", "row_dual", "row_plan2",
                                                  "row_tma", "col_sub", "real_otf", "graph", "k_split"};
    for (size_t i = 0; i < sizeof(kernel_switches) / sizeof(kernel_switches[0]); ++i)
        if (!strcmp(name, kernel_switches[i])) h->e->options_changed();
    if (!strcmp(name, "exact_clip")) {
        h->e->set_exact_clip(value != 0);
        return LSTED_OK;
    }
    if (!strcmp(name, "forget_normalization")) {
        h->e->forget_normalization();
        return LSTED_OK;
    }
    if (!strcmp(name, "reset_estimate")) {
        h->e->reset_estimate();
        return LSTED_OK;
    }
    if (!strcmp(name, "prefetch")) { h->bk->set_prefetch(value != 0); return LSTED_OK; }
    if (!strcmp(name, "fast_path")) { h->bk->set_fast_path(value != 0); return LSTED_OK; }
    if (!strcmp(name, "row_dual")) { h->bk->set_row_dual(value != 0); return LSTED_OK; }
    if (!strcmp(name, "row_plan2")) { h->bk->set_row_plan2(value != 0); return LSTED_OK; }
    if (!strcmp(name, "row_tma")) { h->bk->set_row_tma((int)value); return LSTED_OK; }
    if (!strcmp(name, "col_sub")) { h->bk->set_col_sub(value != 0); return LSTED_OK; }
    if (!strcmp(name, "real_otf")) { h->bk->set_real_otf(value != 0); return LSTED_OK; }   // before set_psfs
    if (!strcmp(name, "profile")) { h->bk->set_profile(value != 0); return LSTED_OK; }
    if (!strcmp(name, "graph")) { h->bk->set_graph(value != 0); return LSTED_OK; }
    if (!strcmp(name, "nvls_ctas")) { h->bk->set_nvls_shape((int)value, 0); return LSTED_OK; }      // before the first iteration
    if (!strcmp(name, "k_split")) { h->bk->set_k_split(value != 0); return LSTED_OK; }
    if (!strcmp(name, "nvls_probe")) { h->bk->set_nvls_probe(value != 0); return LSTED_OK; }
    if (!strcmp(name, "nvls_threads")) { h->bk->set_nvls_shape(-1, (int)value); return LSTED_OK; }
    return set_error(LSTED_ERR_ARG, std::string("unknown option ") + name);
    LSTED_CATCH
}

extern "C" int lsted_deconv_create_data(lsted_deconv* h, const double* object, double total_brightness,
                                        int rescale, uint64_t seed) {
    if (!h || !object) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    LSTED_ENGINE(h, create_data(object, total_brightness, rescale != 0, seed));
    h->bk->sync();
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_upload_object(lsted_deconv* h, const double* object) {
    if (!h || !object) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    LSTED_ENGINE(h, upload_object(object));
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_simulate(lsted_deconv* h, double total_brightness, int rescale,
                                     uint64_t seed) {
    if (!h) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    LSTED_ENGINE(h, simulate(total_brightness, rescale != 0, seed));
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_shard(lsted_deconv* h, int rank, int world, int k_offset,
                                  const char* unique_id) {
    if (!h) return set_error(LSTED_ERR_ARG, "null pointer");
    if (world < 1 || rank < 0 || rank >= world || k_offset < 0)
        return set_error(LSTED_ERR_ARG, "bad rank / world / k_offset");
    LSTED_TRY
    h->bk->activate();
    if (world > 1) h->bk->comm_init(unique_id, rank, world);
    h->e->set_sharding(rank, world, k_offset);
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_p2p_export(lsted_deconv* h, char* handles_out, int capacity) {
    if (!h || !handles_out) return set_error(LSTED_ERR_ARG, "null pointer");
    if (capacity < LSTED_P2P_HANDLE_BYTES) return set_error(LSTED_ERR_ARG, "handle buffer too small");
    LSTED_TRY
    h->bk->activate();
    h->e->p2p_export(handles_out);
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_p2p_attach(lsted_deconv* h, const char* all_handles, int world) {
    if (!h || !all_handles || world < 2) return set_error(LSTED_ERR_ARG, "bad arguments");
    LSTED_TRY
    h->bk->activate();
    h->e->p2p_attach(all_handles);
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_nvls_supported(int device, int* supported) {
    if (!supported) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    *supported = LSTED_BACKEND::nvls_device_supported(device) ? 1 : 0;
    return LSTED_OK;
    LSTED_CATCH
}
extern "C" int lsted_deconv_nvls_create(lsted_deconv* h, int world, int* fd_out) {
    if (!h || !fd_out || world < 2) return set_error(LSTED_ERR_ARG, "bad arguments");
    LSTED_TRY
    h->bk->activate();
    *fd_out = h->e->nvls_create(world);
    return LSTED_OK;
    LSTED_CATCH
}
extern "C" int lsted_deconv_nvls_import(lsted_deconv* h, int world, int fd) {
    if (!h || world < 2 || fd < 0) return set_error(LSTED_ERR_ARG, "bad arguments");
    LSTED_TRY
    h->bk->activate();
    h->e->nvls_import(world, fd);
    return LSTED_OK;
    LSTED_CATCH
}
extern "C" int lsted_deconv_nvls_add_device(lsted_deconv* h) {
    if (!h) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    h->bk->activate();
    h->e->nvls_add_device();
    return LSTED_OK;
    LSTED_CATCH
}
extern "C" int lsted_deconv_nvls_bind(lsted_deconv* h) {
    if (!h) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    h->bk->activate();
    h->e->nvls_bind();
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_iterate(lsted_deconv* h, int n) {
    if (!h) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    LSTED_ENGINE(h, iterate(n));
    return LSTED_OK;
    LSTED_CATCH
}

static int check_which(lsted_deconv* h, int which, int k) {
    lsted::EngineInfo ei;
    h->e->info(&ei);
    const int K = ei.K;
    if (which < 0 || which > 4) return 1;
    if ((which == LSTED_NOISELESS || which == LSTED_NOISY) ? (k < 0 || k >= K) : (k != 0)) return 1;
    return 0;
}

extern "C" int lsted_deconv_get(lsted_deconv* h, int which, int k, double* out) {
    if (!h || !out) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    if (check_which(h, which, k)) return set_error(LSTED_ERR_ARG, "bad array selector");
    LSTED_ENGINE(h, get_array(which, k, out));
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_set(lsted_deconv* h, int which, int k, const double* in) {
    if (!h || !in) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    if (check_which(h, which, k)) return set_error(LSTED_ERR_ARG, "bad array selector");
    LSTED_ENGINE(h, set_array(which, k, in));
    h->bk->sync();
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_ft_error(lsted_deconv* h, const double* image, double* out, int* done) {
    if (!h || !out || !done) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    h->bk->activate();
    *done = h->e->ft_error(image, out) ? 1 : 0;
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_H(lsted_deconv* h, const double* x, double* out) {
    if (!h || !x || !out) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    LSTED_ENGINE(h, H_host(x, out));
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_Ht(lsted_deconv* h, const double* y, double* out, int normalize) {
    if (!h || !y || !out) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    LSTED_ENGINE(h, Ht_host(y, out, normalize != 0));
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_sync(lsted_deconv* h) {
    if (!h) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    h->bk->activate();
    h->bk->sync();
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_timer_start(lsted_deconv* h) {
    if (!h) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    h->bk->activate();
    h->bk->timer_start();
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_timer_stop(lsted_deconv* h, float* ms) {
    if (!h || !ms) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    h->bk->activate();
    *ms = h->bk->timer_stop();
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_deconv_profile(lsted_deconv* h, int reset, double* total_ms, long long* launches) {
    if (!h) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    h->bk->activate();
    h->bk->profile_collect(total_ms, launches);
    if (reset) h->bk->profile_reset();
    return LSTED_OK;
    LSTED_CATCH
}
