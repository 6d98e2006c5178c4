// In-register FFT building blocks for the hand-written transforms of the Stolt pipeline (stolt_fft.cu).
// Everything here is fully unrolled at compile time: arrays live in registers, twiddles of the small
// DFTs are literals, and the digit permutations are register renames.
#pragma once
#include <cuda_runtime.h>

namespace impdar {
namespace fftr {

typedef float2 cf;

#define FFTR_DI __device__ __forceinline__

FFTR_DI cf mk(float x, float y) { return make_float2(x, y); }
// Complex arithmetic on the packed fp32x2 pipe of sm_100 (FADD2 / FMUL2 / FFMA2): an add is one instruction,
// a multiply two; negations, the lane swap of a multiplication by +-i and scalar broadcasts fold into operand
// modifiers (checked in SASS: R.F32x2.LO_HI.NP, R.F32).
FFTR_DI cf cadd(cf a, cf b) { return __fadd2_rn(a, b); }
FFTR_DI cf csub(cf a, cf b) { return __fadd2_rn(a, mk(-b.x, -b.y)); }
FFTR_DI cf cmul(cf a, cf b) { return __ffma2_rn(a, mk(b.x, b.x), __fmul2_rn(mk(-a.y, a.x), mk(b.y, b.y))); }
FFTR_DI cf cmul_conj(cf a, cf b) { return __ffma2_rn(a, mk(b.x, b.x), __fmul2_rn(mk(a.y, -a.x), mk(b.y, b.y))); }  // a * conj(b)
FFTR_DI cf cconj(cf a) { return mk(a.x, -a.y); }
FFTR_DI cf cscale(cf a, float s) { return __fmul2_rn(a, mk(s, s)); }
FFTR_DI cf cmuli(cf a) { return mk(-a.y, a.x); }    // a * i
FFTR_DI cf cmulni(cf a) { return mk(a.y, -a.x); }   // a * (-i)
// s0 * a + s1 * b with real scalars
FFTR_DI cf clerp(cf a, float s0, cf b, float s1) { return __ffma2_rn(a, mk(s0, s0), __fmul2_rn(b, mk(s1, s1))); }
// a * w for the forward transform (DIR < 0), a * conj(w) for the inverse; tables hold forward twiddles e^{-i theta}
template <int DIR>
FFTR_DI cf cmul_dir(cf a, cf w) {
    return DIR < 0 ? cmul(a, w) : cmul_conj(a, w);
}

// cos / sin of 2 pi m / 32 as compile-time constants
__host__ __device__ constexpr float c32q(int m) {
    return m == 0 ? 1.f
         : m == 1 ? 0.98078528040323044913f
         : m == 2 ? 0.92387953251128675613f
         : m == 3 ? 0.83146961230254523708f
         : m == 4 ? 0.70710678118654752440f
         : m == 5 ? 0.55557023301960222474f
         : m == 6 ? 0.38268343236508977173f
         : m == 7 ? 0.19509032201612826785f
                  : 0.f;
}
__host__ __device__ constexpr float cos32(int m) {
    m &= 31;
    return m <= 8 ? c32q(m) : m <= 16 ? -c32q(16 - m) : m <= 24 ? -c32q(m - 16) : c32q(32 - m);
}
__host__ __device__ constexpr float sin32(int m) { return cos32(m + 24); }  // sin(x) = cos(x - pi/2)

// a * e^{DIR * 2 pi i m32 / 32}; m32 is a constant after unrolling, so the branches fold away
template <int DIR>
FFTR_DI cf twmul32(cf a, int m32) {
    m32 &= 31;
    if (m32 == 0) return a;
    if (m32 == 8) return DIR < 0 ? mk(a.y, -a.x) : mk(-a.y, a.x);
    if (m32 == 16) return mk(-a.x, -a.y);
    if (m32 == 24) return DIR < 0 ? mk(-a.y, a.x) : mk(a.y, -a.x);
    const float c = cos32(m32);
    const float s = DIR < 0 ? -sin32(m32) : sin32(m32);
    return cmul(a, mk(c, s));
}

template <int DIR>
FFTR_DI void fft4(cf &a0, cf &a1, cf &a2, cf &a3) {
    const cf s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
    const cf r = DIR < 0 ? mk(d13.y, -d13.x) : mk(-d13.y, d13.x);  // -+ i * d13
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cadd(d02, r);
    a3 = csub(d02, r);
}

// DFT of size N (1, 2, 4, 8, 16, 32) on v[0..N), natural order in and out.
template <int N, int DIR>
FFTR_DI void fft_reg(cf *v) {
    if constexpr (N == 1) {
    } else if constexpr (N == 2) {
        const cf a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else if constexpr (N == 4) {
        fft4<DIR>(v[0], v[1], v[2], v[3]);
    } else {
        // n = a * B + b (a < 4), k = k1 + 4 * k2: X[k] = sum_b w_N^{b k1} w_B^{b k2} sum_a x[aB + b] w_4^{a k1}
        constexpr int B = N / 4;
        cf y[4][B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            cf t0 = v[b], t1 = v[B + b], t2 = v[2 * B + b], t3 = v[3 * B + b];
            fft4<DIR>(t0, t1, t2, t3);
            y[0][b] = t0;
            y[1][b] = twmul32<DIR>(t1, b * (32 / N));
            y[2][b] = twmul32<DIR>(t2, 2 * b * (32 / N));
            y[3][b] = twmul32<DIR>(t3, 3 * b * (32 / N));
        }
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            fft_reg<B, DIR>(y[k1]);
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) v[k1 + 4 * k2] = y[k1][k2];
        }
    }
}

// v[t] *= w^t (forward twiddle w; the inverse uses conj(w)).  Four interleaved chains c_r = w^r (w^4)^m keep only
// five twiddles live (a full power table would double the register footprint of a radix-32 item) and bound the
// rounding error growth at R/4 + 2 multiplications.
template <int R, int DIR>
FFTR_DI void apply_powers(cf *v, cf w) {
    if (DIR > 0) w.y = -w.y;
    if constexpr (R <= 4) {
        cf p = w;
#pragma unroll
        for (int t = 1; t < R; ++t) {
            v[t] = cmul(v[t], p);
            if (t + 1 < R) p = cmul(p, w);
        }
    } else {
        const cf w2 = cmul(w, w);
        const cf w4 = cmul(w2, w2);
        cf c[4];
        c[0] = w4;  // first used at t = 4
        c[1] = w;
        c[2] = w2;
        c[3] = cmul(w2, w);
#pragma unroll
        for (int t = 1; t < R; ++t) {
            v[t] = cmul(v[t], c[t & 3]);
            if (t + 4 < R) c[t & 3] = cmul(c[t & 3], w4);
        }
    }
}

}  // namespace fftr
}  // namespace impdar
