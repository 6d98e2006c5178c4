// Transposed direct-form-II IIR recurrence shared by the filtfilt kernels (scipy.signal.lfilter semantics).
#pragma once

namespace impdar {

struct IirCoef {
    double b[33];
    double a[33];
    double zi[32];
};

// One step of the transposed direct-form II recurrence (scipy.signal.lfilter).  The feed-forward halves
// t_i = b_{i+1} x + z_{i+1} do not depend on y, so the serial chain per sample is two DFMA (y, then z_0).
// BZ: every odd-indexed b is exactly zero (any band-pass design from scipy: the numerator is k (1 - z^-2)^n), so
// those feed-forward DFMAs are dropped - fma(0, x, z) == z bit for bit for finite x.
template <typename T, int NS, bool BZ = false>
__device__ __forceinline__ double iir_step(double xv, double (&z)[NS], const IirCoef &c) {
    const double yv = fma(c.b[0], xv, z[0]);
#pragma unroll
    for (int i = 0; i < NS - 1; ++i) {
        if (BZ && ((i + 1) & 1)) z[i] = fma(-c.a[i + 1], yv, z[i + 1]);
        else z[i] = fma(-c.a[i + 1], yv, fma(c.b[i + 1], xv, z[i + 1]));
    }
    if (BZ && (NS & 1)) z[NS - 1] = -c.a[NS] * yv;
    else z[NS - 1] = fma(-c.a[NS], yv, c.b[NS] * xv);
    return yv;
}

}  // namespace impdar
