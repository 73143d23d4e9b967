// Small elementwise / reduction bodies shared by the CUDA and the
// host-replay backends (grid-stride over `i`).
#pragma once
#include "fft_core.cuh"

namespace lsted {

enum EwOp { EW_CAST_IN = 0, EW_CAST_OUT = 1, EW_FILL = 2, EW_DIVIDE = 3, EW_RL_UPDATE = 4 };

template <typename T> struct EwArgs {
    T* t0;            // destination (or in/out)
    const T* t1;
    const T* t2;
    double* d0;
    const double* d1;
    double s;
    size_t n;
};

template <int OP, typename T> LSTED_HD void ew_apply(const EwArgs<T>& a, size_t i) {
    if (OP == EW_CAST_IN) a.t0[i] = (T)(a.d1[i] * a.s);
    else if (OP == EW_CAST_OUT) a.d0[i] = (double)a.t1[i];
    else if (OP == EW_FILL) a.t0[i] = (T)a.s;
    else if (OP == EW_DIVIDE) a.t0[i] = a.t0[i] / a.t1[i];
    else if (OP == EW_RL_UPDATE) a.t0[i] = a.t0[i] * (a.t1[i] / a.t2[i]);  // est *= H_t(ratio)/norm
}

}  // namespace lsted
