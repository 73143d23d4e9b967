// extern "C" entry points of the figure-3 scan-position engine and of the general spline /
// Gaussian plane operators (include/lsted.h), written once over the backend type
// LSTED_SCAN_BACKEND (ScanCudaBackend in lsted_api.cu; the CPU replay in tests/host_emul/emul.cpp,
// test infrastructure only).  The including file provides set_error(), LSTED_TRY / LSTED_CATCH.

struct lsted_scan {
    LSTED_SCAN_BACKEND* bk;
    lsted::ScanEngine<LSTED_SCAN_BACKEND>* e;
};

extern "C" int lsted_scan_destroy(lsted_scan* h) {
    if (!h) return LSTED_OK;
    LSTED_TRY
    if (h->bk) h->bk->activate();
    delete h->e;
    delete h->bk;
    delete h;
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_scan_create(lsted_scan** out, int device, const lsted_scan_params_t* p,
                                 const int* positions, const double* blur_taps,
                                 const double* exc_taps) {
    if (!out || !p || !positions || !blur_taps || !exc_taps) return set_error(LSTED_ERR_ARG, "null pointer");
    if (p->imaging_type < 0 || p->imaging_type > 3) return set_error(LSTED_ERR_ARG, "unknown imaging type");
    if (p->n_y < 2 || p->n_x < 1 || p->pad < 1 || p->step < 1 || p->num_positions < 1 ||
        p->blur_radius < 0 || p->exc_radius < 0)
        return set_error(LSTED_ERR_ARG, "bad scan arguments");
    if (p->imaging_type == lsted::SCAN_MULTIPOINT && (p->exc_sep < p->step || p->exc_sep % p->step))
        return set_error(LSTED_ERR_ARG, "spot separation must be a multiple of the scan step");
    if (p->imaging_type == lsted::SCAN_RESCAN_LINE && !(p->zoom_factor > 0.0 && p->zoom_factor <= 1.0))
        return set_error(LSTED_ERR_ARG, "rescan zoom factor must be in (0, 1]");
    lsted_scan* h = new lsted_scan();
    h->bk = 0; h->e = 0;
    try {
        h->bk = new LSTED_SCAN_BACKEND(device);
        h->bk->activate();
        lsted::ScanParams sp;
        sp.type = p->imaging_type;
        sp.n_y = p->n_y; sp.n_x = p->n_x; sp.pad = p->pad;
        sp.step = p->step; sp.exc_sep = p->exc_sep; sp.num_pos = p->num_positions;
        sp.zoom_factor = p->zoom_factor;
        sp.blur_radius = p->blur_radius; sp.exc_radius = p->exc_radius;
        sp.chunk_bytes = p->chunk_bytes ? p->chunk_bytes : ((size_t)4 << 30);
        h->e = new lsted::ScanEngine<LSTED_SCAN_BACKEND>(*h->bk, sp, positions, blur_taps, exc_taps);
        h->bk->sync();
        *out = h;
        return LSTED_OK;
    }
    catch (const lsted::ApiError& e) { lsted_scan_destroy(h); return set_error(e.code, e.msg); }
    catch (const std::string& s) { lsted_scan_destroy(h); return set_error(LSTED_ERR_ARG, s); }
    catch (const std::bad_alloc&) { lsted_scan_destroy(h); return set_error(LSTED_ERR_ARG, "host allocation failed"); }
    catch (...) { lsted_scan_destroy(h); return set_error(LSTED_ERR_ARG, "unexpected exception in lsted_scan_create"); }
}

extern "C" int lsted_scan_excitation(lsted_scan* h, double* out) {
    if (!h || !out) return set_error(LSTED_ERR_ARG, "null pointer");
    LSTED_TRY
    h->bk->activate();
    h->e->get_excitation(out);
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_scan_run(lsted_scan* h, const double* obj_padded, const double* rot_xform,
                              const int* frame_positions, int num_frames, double* maxima,
                              double* reconstruction, double* cum_detector_sig, double* device_ms) {
    if (!h || !obj_padded) return set_error(LSTED_ERR_ARG, "null pointer");
    if (num_frames < 0 || (num_frames && !frame_positions)) return set_error(LSTED_ERR_ARG, "bad frame list");
    for (int f = 0; f < num_frames; ++f)
        if (frame_positions[f] < 0 || frame_positions[f] >= h->e->g.num_pos ||
            (f && frame_positions[f] <= frame_positions[f - 1]))
            return set_error(LSTED_ERR_ARG, "frame positions must be ascending scan-position indices");
    LSTED_TRY
    h->bk->activate();
    h->e->run(obj_padded, rot_xform, frame_positions, num_frames, maxima, reconstruction, cum_detector_sig);
    if (device_ms) *device_ms = h->e->last_ms;
    return LSTED_OK;
    LSTED_CATCH
}

extern "C" int lsted_scan_frames(lsted_scan* h, int first, int count, const double* display_max,
                                 const double* inv_xform, double* out) {
    if (!h || !display_max || !out) return set_error(LSTED_ERR_ARG, "null pointer");
    if (!h->e->have_run) return set_error(LSTED_ERR_STATE, "lsted_scan_frames before lsted_scan_run");
    if (first < 0 || count < 1 || first + count > (int)h->e->h_frame_pos.size())
        return set_error(LSTED_ERR_ARG, "frame range outside the kept frames of the last run");
    LSTED_TRY
    h->bk->activate();
    h->e->frames(first, count, display_max, inv_xform, out);
    return LSTED_OK;
    LSTED_CATCH
}

// scipy.ndimage.affine_transform / rotate / shift / zoom for order 3 on a batch of planes
extern "C" int lsted_img_spline(int device, int batch, int n0, int n1, const double* planes,
                                const double* xform, int m0, int m1, int mode, int clip, double* out) {
    if (!planes || !xform || !out) return set_error(LSTED_ERR_ARG, "null pointer");
    if (batch < 1 || n0 < 1 || n1 < 1 || m0 < 1 || m1 < 1 ||
        (mode != lsted::SPLINE_CONSTANT && mode != lsted::SPLINE_NEAREST))
        return set_error(LSTED_ERR_ARG, "bad spline arguments");
    LSTED_TRY
    LSTED_SCAN_BACKEND bk(device);
    bk.activate();
    const int npad = mode == lsted::SPLINE_NEAREST ? lsted::kSplinePrepad : 0;
    const int c0 = n0 + 2 * npad, c1 = n1 + 2 * npad;
    const size_t in_elems = (size_t)batch * n0 * n1, co_elems = (size_t)batch * c0 * c1;
    const size_t out_elems = (size_t)batch * m0 * m1;
    double* d_in = bk.alloc<double>(in_elems);
    double* d_co = npad ? bk.alloc<double>(co_elems) : d_in;
    double* d_out = bk.alloc<double>(out_elems);
    double* d_xf = bk.alloc<double>(6 * (size_t)batch);
    double* d_clip = clip ? bk.alloc<double>(batch) : nullptr;
    double* d_part = clip ? bk.alloc<double>((size_t)batch * lsted::kReduceSegments) : nullptr;
    bk.upload(d_in, planes, sizeof(double) * in_elems);
    bk.upload(d_xf, xform, sizeof(double) * 6 * batch);
    if (clip) {
        lsted::PlaneReducePartialFn a{d_in, d_part, (size_t)n0 * n1, 1};
        bk.for_each((size_t)batch * lsted::kReduceSegments, a);
        lsted::PlaneReduceFinalFn b{d_part, d_clip, 1, 1, 1.1};
        bk.for_each(batch, b);
    }
    if (npad) {
        lsted::ImgEdgePadFn pad{d_in, d_co, n0, n1, npad};
        bk.for_each(co_elems, pad);
    }
    {   // prefilter along both axes: blocked passes (second buffer) for long lines
        double* d_tmp = bk.alloc<double>(co_elems);
        for (int axis = 0; axis < 2; ++axis) {
            const int n = axis == 0 ? c0 : c1;
            const size_t nlines = axis == 0 ? c1 : c0;
            if (n < lsted::kPfMinLine) {
                lsted::ImgPrefilterFn pf{d_co, c0, c1, axis, npad ? 1 : 0};
                bk.for_each((size_t)batch * nlines, pf);
                continue;
            }
            lsted::PrefilterLines L;
            L.n = n; L.stride = axis == 0 ? (size_t)c1 : 1; L.nlines = nlines;
            L.line_step = axis == 0 ? 1 : (size_t)c1; L.plane_step = (size_t)c0 * c1; L.offset = 0;
            L.nblocks = (n + lsted::kPfBlock - 1) / lsted::kPfBlock;
            lsted::SplineCausalBlockFn ca{d_co, d_tmp, L, (size_t)batch * nlines, npad ? 1 : 0};
            bk.for_each((size_t)batch * nlines * L.nblocks, ca);
            lsted::SplineAnticausalBlockFn an{d_tmp, d_co, L, (size_t)batch * nlines, npad ? 1 : 0};
            bk.for_each((size_t)batch * nlines * L.nblocks, an);
        }
        bk.free(d_tmp);
    }
    lsted::ImgSplineFn sp{d_co, d_xf, d_clip, d_out, c0, c1, m0, m1, npad, mode};
    bk.for_each(out_elems, sp);
    bk.download(out, d_out, sizeof(double) * out_elems);
    bk.free(d_in); bk.free(d_out); bk.free(d_xf);
    if (npad) bk.free(d_co);
    if (clip) { bk.free(d_clip); bk.free(d_part); }
    return LSTED_OK;
    LSTED_CATCH
}

// scipy.ndimage.gaussian_filter (mode='reflect') along the two plane axes; a null tap pointer
// skips that axis (sigma 0)
extern "C" int lsted_img_gauss(int device, int batch, int n0, int n1, const double* planes,
                               const double* taps0, int radius0, const double* taps1, int radius1,
                               double* out) {
    if (!planes || !out) return set_error(LSTED_ERR_ARG, "null pointer");
    if (batch < 1 || n0 < 1 || n1 < 1 || radius0 < 0 || radius1 < 0)
        return set_error(LSTED_ERR_ARG, "bad filter arguments");
    LSTED_TRY
    LSTED_SCAN_BACKEND bk(device);
    bk.activate();
    const size_t elems = (size_t)batch * n0 * n1;
    double* d_a = bk.alloc<double>(elems);
    double* d_b = bk.alloc<double>(elems);
    bk.upload(d_a, planes, sizeof(double) * elems);
    double *src = d_a, *dst = d_b;
    for (int axis = 0; axis < 2; ++axis) {
        const double* taps = axis ? taps1 : taps0;
        const int radius = axis ? radius1 : radius0;
        if (!taps) continue;
        double* d_t = bk.alloc<double>(2 * radius + 1);
        bk.upload(d_t, taps, sizeof(double) * (2 * radius + 1));
        lsted::ImgFirFn f{src, dst, n0, n1, axis, radius, d_t};
        bk.for_each(elems, f);
        bk.free(d_t);
        std::swap(src, dst);
    }
    bk.download(out, src, sizeof(double) * elems);
    bk.free(d_a); bk.free(d_b);
    return LSTED_OK;
    LSTED_CATCH
}
