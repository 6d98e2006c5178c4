// Gazdag phase-shift migration, constant and layered velocity
// (reference: migrationlib/mig_python.py:211-287 migrationPhaseShift, :361-493 phaseShift).
//
//   FK = fft2(taper(data), (nt, tnum)), nt = next pow2 >= snum
//   const v : TK[tau,k] = 1/snum * sum_w [ (v kx/2)^2 < w^2 ] FK[w,k] * cp(w,k)^(tau+1),
//             cp = exp(+i w dt sqrt(1 - (v kx/2)^2 / w^2))                                  (:403-420)
//   v(tau)  : per tau, FK[w,k] *= exp(+i w dt Re sqrt(coss)), coss = 1 - (v_tau kx / 2w)^2;
//             FK[w, coss <= thr2_tau] = 0 (sticky); TK[tau,k] = 1/snum * sum_w FK[w,k]      (:439-487)
//   out = ifft_k(TK).real                                                                   (:282)
//
// Data are real, so FK(-w,-k) = conj FK(w,k) and TK(tau,-k) = conj TK(tau,k): only k = 0..tnum/2 is
// computed (R2C along traces, C2C along time, C2R back), halving the work.  One bin breaks the symmetry:
// the Nyquist frequency (fftfreq's -nt/2) has no +nt/2 partner, so TK is not exactly Hermitian and the
// reference's ``.real`` symmetrises it; for that bin this means FK * Re(cp^(tau+1)), i.e. the cosine only.
//
// The contraction over w is a per-kx non-uniform DFT (the "matrix" depends on kx), i.e. there is no
// operand shared between columns, so it is kept as complex FMA on the SIMT pipes:
//   * a CTA owns 2 kx columns and all nt frequencies; a thread keeps 16 (w,k) states in registers;
//   * const v: the reference's own recurrence FFK *= cp, re-seeded every 64 steps from G = FK*cp^(64a)
//     (G and cp^64 live in shared memory, both built from fp64 phases) so fp32 round-off cannot
//     random-walk over thousands of steps;
//   * layered: the cumulative phase is carried in fp64 turns (sqrt by fp32 rsqrt + one fp64 Newton step)
//     and applied to the original FK with one fp32 sincos per (tau,w,k);
//   * per-tau sums over w: in-thread, xor-shuffle, then one shared-memory pass per 64 taus.
#include <cufft.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

extern "C" int impdar_taper_f32(const float *x, float *y, int S, int T, int batch, double htaper, double vtaper,
                                int trunc_int, void *stream);

namespace impdar {

// phaseshift_tc.cu: constant velocity as a per-kx complex matrix product on the tensor cores
size_t phsh_tc_workspace_bytes(int nt, int K);
int phsh_const_tc_launch(const float2 *FK, float2 *TK, int nt, int K, int S, int T, double dt, double dx, double vel,
                         float inv_s, void *ws, cudaStream_t st);

#define IMPDAR_CUFFT(call)                                                                  \
    do {                                                                                    \
        cufftResult r__ = (call);                                                           \
        if (r__ != CUFFT_SUCCESS) {                                                         \
            impdar::set_error("%s:%d %s -> cufft error %d", __FILE__, __LINE__, #call, (int)r__); \
            return IMPDAR_B200_ECUFFT;                                                      \
        }                                                                                   \
    } while (0)

constexpr int PS_THREADS = 512;
constexpr int PS_TB = 64;  // taus per block (re-seed interval of the constant-velocity recurrence)

struct PhshParams {
    const float2 *FK;  // (nt, K)
    float2 *TK;        // (S, K)
    int nt, K, S, T;
    double dt, dx, vel;
    const double *vmig;  // S (layered) or null
    const double *thr2;  // S (layered) or null
    float inv_s;
};

// ws = 2.*np.pi*np.fft.fftfreq(nt, d=dt) and kx = 2.*np.pi*np.fft.fftfreq(tnum, d=dx) (:268-269) with numpy's OWN float64
// operation sequence - fftfreq is `results * (1.0 / (n * d))` on integer results, then one multiply by the double
// 2*pi - each step rounded on its own.  This matters: the propagating / evanescent decision `vkx2 < w**2.` (:411-412)
// is an EXACT TIE for whole families of (w, kx) bins on the usual "nice" geometries (dt = 1e-8, dx = 5, v = 1.69e8,
// power-of-two sizes: v k nt dt / (2 tnum dx) is an integer), and such a bin - phase ~ 0 - adds a tau-independent
// FK / (snum tnum) to TK, i.e. 1e-3 of relative L2, if it is classified differently from the reference.
#define PS_TWO_PI 6.283185307179586
__device__ __forceinline__ double ps_omega(int iw, int nt, double dt) {
    // with w == 0 replaced by 1e-10/dt (:404-406, :446-448)
    const int fi = (iw < (nt + 1) / 2) ? iw : iw - nt;
    if (fi == 0) return 1e-10 / dt;
    const double val = __ddiv_rn(1.0, __dmul_rn((double)nt, dt));
    return __dmul_rn(PS_TWO_PI, __dmul_rn((double)fi, val));
}
__device__ __forceinline__ double ps_kx(int k, int T, double dx) {
    const double val = __ddiv_rn(1.0, __dmul_rn((double)T, dx));
    return __dmul_rn(PS_TWO_PI, __dmul_rn((double)k, val));  // k <= T/2: the non-negative branch (kx enters squared)
}
// (vmig*kx/2.)**2. (:411), w**2., and -phase = w*dt*sqrt(1.0 - vkx2/w**2.) (:415), every operation rounded as numpy does
__device__ __forceinline__ double ps_vkx2(double vel, double kx) {
    const double h = __ddiv_rn(__dmul_rn(vel, kx), 2.0);
    return __dmul_rn(h, h);
}
__device__ __forceinline__ double ps_phi(double w, double dt, double vkx2, double w2) {
    return __dmul_rn(__dmul_rn(w, dt), __dsqrt_rn(__dsub_rn(1.0, __ddiv_rn(vkx2, w2))));
}
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// Sum `part` over the lanes that share a kx column and park the warp's result for tau slot b.
template <int COLS>
__device__ __forceinline__ void ps_park(float2 part, float2 (*buf)[PS_THREADS / 32][COLS], int b) {
#pragma unroll
    for (int o = COLS; o < 32; o <<= 1) {
        part.x += __shfl_xor_sync(0xffffffffu, part.x, o);
        part.y += __shfl_xor_sync(0xffffffffu, part.y, o);
    }
    const int lane = threadIdx.x & 31;
    if (lane < COLS) buf[b][threadIdx.x >> 5][lane] = part;
}

// Sum the warps' parked partials of one tau block and store (or accumulate) TK.
// `nyq_const` adds the constant-velocity Nyquist-frequency term FK[nt/2,k] * cos((tau+1) phi) (see header).
template <int COLS>
__device__ __forceinline__ void ps_commit(const PhshParams &p, float2 (*buf)[PS_THREADS / 32][COLS], int tb0,
                                          bool accumulate, bool nyq_const = false) {
    __syncthreads();
    for (int i = threadIdx.x; i < PS_TB * COLS; i += PS_THREADS) {
        const int b = i / COLS, c = i % COLS;
        const int tau = tb0 + b, kk = blockIdx.x * COLS + c;
        if (tau < p.S && kk < p.K) {
            float2 s = make_float2(0.f, 0.f);
#pragma unroll
            for (int w = 0; w < PS_THREADS / 32; ++w) {
                s.x += buf[b][w][c].x;
                s.y += buf[b][w][c].y;
            }
            if (nyq_const && p.nt >= 2) {
                const double w = ps_omega(p.nt / 2, p.nt, p.dt);
                const double vkx2n = ps_vkx2(p.vel, ps_kx(kk, p.T, p.dx)), w2n = __dmul_rn(w, w);
                if (vkx2n < w2n) {
                    const double phi = ps_phi(w, p.dt, vkx2n, w2n);
                    const float cn = (float)cos((double)(tau + 1) * phi);
                    const float2 f = p.FK[(size_t)(p.nt / 2) * p.K + kk];
                    s.x = fmaf(f.x, cn, s.x);
                    s.y = fmaf(f.y, cn, s.y);
                }
            }
            s.x *= p.inv_s;
            s.y *= p.inv_s;
            float2 *dst = p.TK + (size_t)tau * p.K + kk;
            if (accumulate) {
                const float2 o = *dst;
                s.x += o.x;
                s.y += o.y;
            }
            *dst = s;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------ constant velocity (:396-420)
constexpr int PSC_COLS = 2;
constexpr int PSC_PER = 16;
constexpr int PSC_WCHUNK = PS_THREADS / PSC_COLS * PSC_PER;  // 4096 frequencies per pass

__global__ void __launch_bounds__(PS_THREADS, 1) phsh_const_kernel(const __grid_constant__ PhshParams p) {
    extern __shared__ float2 ps_smem[];
    float2 *G = ps_smem;                          // [PSC_PER][PS_THREADS]  FK * cp^(tb0)
    float2 *Z64 = ps_smem + PSC_PER * PS_THREADS;  // [PSC_PER][PS_THREADS]  cp^64
    __shared__ float2 buf[PS_TB][PS_THREADS / 32][PSC_COLS];

    const int tid = threadIdx.x;
    const int k = blockIdx.x * PSC_COLS + (tid % PSC_COLS);
    const bool kvalid = k < p.K;
    const double vkx2 = ps_vkx2(p.vel, ps_kx(kvalid ? k : 0, p.T, p.dx));  // (vmig*kx/2)^2   (:411)

    for (int ch = 0; ch * PSC_WCHUNK < p.nt; ++ch) {
        float2 z[PSC_PER];
#pragma unroll
        for (int i = 0; i < PSC_PER; ++i) {
            const int iw = ch * PSC_WCHUNK + (tid / PSC_COLS) + i * (PS_THREADS / PSC_COLS);
            float2 g = make_float2(0.f, 0.f), zz = make_float2(1.f, 0.f), z64 = zz;
            if (kvalid && iw < p.nt && !(p.nt >= 2 && iw == p.nt / 2)) {  // the Nyquist bin is added at commit
                const double w = ps_omega(iw, p.nt, p.dt);
                const double w2 = __dmul_rn(w, w);
                if (vkx2 < w2) {  // propagating (:412)
                    const double phi = ps_phi(w, p.dt, vkx2, w2);  // = -phase (:415); cp = e^{+i phi}
                    double sn, cs;
                    sincos(phi, &sn, &cs);
                    zz = make_float2((float)cs, (float)sn);
                    sincos(phi * (double)PS_TB, &sn, &cs);
                    z64 = make_float2((float)cs, (float)sn);
                    g = p.FK[(size_t)iw * p.K + k];
                }
            }
            z[i] = zz;
            G[i * PS_THREADS + tid] = g;
            Z64[i * PS_THREADS + tid] = z64;
        }
        for (int tb0 = 0; tb0 < p.S; tb0 += PS_TB) {
            float2 ffk[PSC_PER];
#pragma unroll
            for (int i = 0; i < PSC_PER; ++i) ffk[i] = G[i * PS_THREADS + tid];
            const int nb = min(PS_TB, p.S - tb0);
            for (int b = 0; b < nb; ++b) {
                float2 part = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < PSC_PER; ++i) {
                    ffk[i] = cmulf(ffk[i], z[i]);  // FFK *= cp      (:419)
                    part.x += ffk[i].x;            // TK[itau] += FFK (:420)
                    part.y += ffk[i].y;
                }
                ps_park<PSC_COLS>(part, buf, b);
            }
#pragma unroll
            for (int i = 0; i < PSC_PER; ++i)
                G[i * PS_THREADS + tid] = cmulf(G[i * PS_THREADS + tid], Z64[i * PS_THREADS + tid]);
            ps_commit<PSC_COLS>(p, buf, tb0, ch > 0, ch == 0);
        }
    }
}

// --------------------------------------------------------------------------- layered v(tau) (:439-487)
constexpr int PSL_PER = 8;
constexpr int PSL_WCHUNK = PS_THREADS * PSL_PER;  // 4096 frequencies per pass, one kx column per CTA

__global__ void __launch_bounds__(PS_THREADS, 1) phsh_layered_kernel(const __grid_constant__ PhshParams p) {
    __shared__ float2 buf[PS_TB][PS_THREADS / 32][1];
    const int tid = threadIdx.x;
    const int k = blockIdx.x;
    const double kx = ps_kx(k, p.T, p.dx);

    for (int ch = 0; ch * PSL_WCHUNK < p.nt; ++ch) {
        float2 fk0[PSL_PER];
        double c[PSL_PER], wturn[PSL_PER], phase[PSL_PER];
        unsigned nyq_slot = 0;  // bit i set: state i is the Nyquist bin, which contributes FK * cos(phase) only
#pragma unroll
        for (int i = 0; i < PSL_PER; ++i) {
            const int iw = ch * PSL_WCHUNK + tid + i * PS_THREADS;
            fk0[i] = make_float2(0.f, 0.f);
            c[i] = 0.0;
            wturn[i] = 0.0;
            phase[i] = 0.0;
            if (iw < p.nt) {
                const double w = ps_omega(iw, p.nt, p.dt);
                const double h = 0.5 * kx / w;
                c[i] = h * h;                                        // coss = 1 - c * v^2   (:460)
                wturn[i] = w * p.dt * 0.15915494309189535;           // phase advance in turns per unit sqrt(coss)
                fk0[i] = p.FK[(size_t)iw * p.K + k];
                if (p.nt >= 2 && iw == p.nt / 2) nyq_slot |= 1u << i;
            }
        }
        for (int tb0 = 0; tb0 < p.S; tb0 += PS_TB) {
            const int nb = min(PS_TB, p.S - tb0);
            for (int b = 0; b < nb; ++b) {
                const double v = p.vmig[tb0 + b];
                const double v2 = v * v;
                const double th = p.thr2[tb0 + b];
                float2 part = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < PSL_PER; ++i) {
                    const double coss = fma(-c[i], v2, 1.0);
                    if (coss <= th) {  // evanescent from here on (:484-485, sticky because FK itself is zeroed)
                        fk0[i] = make_float2(0.f, 0.f);
                    } else {
                        // sqrt(coss): fp32 rsqrt seed + one fp64 Newton step (error ~1e-14)
                        const float cf = (float)coss;
                        float rs;
                        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(cf));
                        const double s0 = (double)(cf * rs);
                        const double s1 = fma(fma(-s0, s0, coss), (double)(0.5f * rs), s0);
                        phase[i] = fma(wturn[i], s1, phase[i]);
                        const double fr = phase[i] - rint(phase[i]);
                        float sn, cs;
                        __sincosf((float)fr * 6.283185307179586f, &sn, &cs);
                        if ((nyq_slot >> i) & 1u) sn = 0.f;
                        part.x += fmaf(fk0[i].x, cs, -fk0[i].y * sn);
                        part.y += fmaf(fk0[i].x, sn, fk0[i].y * cs);
                    }
                }
                ps_park<1>(part, buf, b);
            }
            ps_commit<1>(p, buf, tb0, ch > 0);
        }
    }
}

// =====================================================================================================
// (+w, -w) pair kernels.  The phase of the reference is odd in w (phi(-w) = -phi(w)) while the propagating /
// evanescent decision depends on w^2 only, so the two bins share one square root, one sincos and one mask:
//     FK[+w] e^{+i phi} + FK[-w] e^{-i phi} = cos(phi) A + i sin(phi) B,   A = FK[+w] + FK[-w], B = FK[+w] - FK[-w].
// That halves the MUFU and fp64 work per migrated sample and replaces two complex multiplies by four FMAs
// (two FFMA2 on the packed fp32x2 pipe).  States are the pairs s = 0 .. nt/2 - 1 (s = 0 is the single bin w = 0,
// A = B = FK[0]); only pairs that can still propagate get a thread slot; the unpaired Nyquist bin
// (cosine only, see the header) is added at commit time (constant velocity) or by a one-thread-per-kx kernel.
constexpr int PP_THREADS = 256;
constexpr int PP_PER = 8;
constexpr int PP_STATES = PP_THREADS * PP_PER;

__device__ __forceinline__ float2 pp_mk(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ float2 pp_cmul(float2 a, float2 b) {
    return __ffma2_rn(a, pp_mk(b.x, b.x), __fmul2_rn(pp_mk(-a.y, a.x), pp_mk(b.y, b.y)));
}
// acc + cs * A + i sn * B
__device__ __forceinline__ float2 pp_acc(float2 acc, float cs, float sn, float2 A, float2 B) {
    return __ffma2_rn(pp_mk(-B.y, B.x), pp_mk(sn, sn), __ffma2_rn(A, pp_mk(cs, cs), acc));
}

template <int NW>
__device__ __forceinline__ void pp_park(float2 part, float2 (*buf)[NW], int b) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        part.x += __shfl_xor_sync(0xffffffffu, part.x, o);
        part.y += __shfl_xor_sync(0xffffffffu, part.y, o);
    }
    if ((threadIdx.x & 31) == 0) buf[b][threadIdx.x >> 5] = part;
}

// Sum the warps' partials of one tau block into TK[tau, k] (scaled); `nyq_const` adds the constant-velocity
// Nyquist term FK[nt/2, k] cos((tau + 1) phi).
template <int NW>
__device__ __forceinline__ void pp_commit(const PhshParams &p, float2 (*buf)[NW], int k, int tb0, bool accumulate,
                                          bool nyq_const) {
    __syncthreads();
    for (int b = threadIdx.x; b < PS_TB; b += blockDim.x) {
        const int tau = tb0 + b;
        if (tau < p.S) {
            float2 s = make_float2(0.f, 0.f);
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                s.x += buf[b][w].x;
                s.y += buf[b][w].y;
            }
            if (nyq_const && p.nt >= 2) {
                const double w = ps_omega(p.nt / 2, p.nt, p.dt);
                const double vkx2n = ps_vkx2(p.vel, ps_kx(k, p.T, p.dx)), w2n = __dmul_rn(w, w);
                if (vkx2n < w2n) {
                    const double phi = ps_phi(w, p.dt, vkx2n, w2n);
                    const float cn = (float)cos((double)(tau + 1) * phi);
                    const float2 f = p.FK[(size_t)(p.nt / 2) * p.K + k];
                    s.x = fmaf(f.x, cn, s.x);
                    s.y = fmaf(f.y, cn, s.y);
                }
            }
            s.x *= p.inv_s;
            s.y *= p.inv_s;
            float2 *dst = p.TK + (size_t)tau * p.K + k;
            if (accumulate) {
                const float2 o = *dst;
                s.x += o.x;
                s.y += o.y;
            }
            *dst = s;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void pp_load_pair(const PhshParams &p, int s, int k, float2 &A, float2 &B) {
    const float2 fp = p.FK[(size_t)s * p.K + k];
    if (s == 0) {
        A = fp;
        B = fp;
    } else {
        const float2 fm = p.FK[(size_t)(p.nt - s) * p.K + k];
        A = make_float2(fp.x + fm.x, fp.y + fm.y);
        B = make_float2(fp.x - fm.x, fp.y - fm.y);
    }
}

// ---- constant velocity (:396-420): rotation recurrence r *= e^{i phi}, re-seeded every PS_TB steps from fp64 seeds
__global__ void __launch_bounds__(PP_THREADS, 2) phsh_const_pair_kernel(const __grid_constant__ PhshParams p) {
    __shared__ float2 Rb[PP_PER][PP_THREADS];    // e^{i phi tb0}
    __shared__ float2 Z64[PP_PER][PP_THREADS];   // e^{i phi PS_TB}
    __shared__ float2 buf[PS_TB][PP_THREADS / 32];
    const int tid = threadIdx.x;
    const int k = blockIdx.x;
    const int nh = max(1, p.nt / 2);  // nt == 1: the single bin w = 0
    const double vkx2 = ps_vkx2(p.vel, ps_kx(k, p.T, p.dx));  // (vmig*kx/2)^2   (:411)
    const double vk = sqrt(vkx2);
    // pairs below s ~ vk nt dt / 2 pi are evanescent for every tau (:412); start two below the estimate (ties are
    // decided by the exact comparison below, never by this estimate)
    int s_first = (int)fmin((double)nh, floor(vk * (double)p.nt * p.dt * 0.15915494309189535)) - 2;
    if (s_first < 0) s_first = 0;

    int pass = 0;
    for (int base = s_first; base < nh || pass == 0; base += PP_STATES, ++pass) {
        float2 z[PP_PER], A[PP_PER], B[PP_PER];
#pragma unroll
        for (int i = 0; i < PP_PER; ++i) {
            const int s = base + tid + i * PP_THREADS;
            z[i] = make_float2(1.f, 0.f);
            A[i] = make_float2(0.f, 0.f);
            B[i] = A[i];
            float2 z64 = z[i];
            if (s < nh) {
                const double w = ps_omega(s, p.nt, p.dt);
                const double w2 = __dmul_rn(w, w);
                if (vkx2 < w2) {  // propagating (:412)
                    const double phi = ps_phi(w, p.dt, vkx2, w2);  // = -phase (:415); cp = e^{+i phi}
                    double sn, cs;
                    sincos(phi, &sn, &cs);
                    z[i] = make_float2((float)cs, (float)sn);
                    sincos(phi * (double)PS_TB, &sn, &cs);
                    z64 = make_float2((float)cs, (float)sn);
                    pp_load_pair(p, s, k, A[i], B[i]);
                }
            }
            Rb[i][tid] = make_float2(1.f, 0.f);
            Z64[i][tid] = z64;
        }
        for (int tb0 = 0; tb0 < p.S; tb0 += PS_TB) {
            float2 r[PP_PER];
#pragma unroll
            for (int i = 0; i < PP_PER; ++i) r[i] = Rb[i][tid];
            const int nb = min(PS_TB, p.S - tb0);
            for (int b = 0; b < nb; ++b) {
                float2 part = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < PP_PER; ++i) {
                    r[i] = pp_cmul(r[i], z[i]);                       // FFK *= cp      (:419)
                    part = pp_acc(part, r[i].x, r[i].y, A[i], B[i]);  // TK[itau] += FFK (:420), both signs of w
                }
                pp_park<PP_THREADS / 32>(part, buf, b);
            }
#pragma unroll
            for (int i = 0; i < PP_PER; ++i) Rb[i][tid] = pp_cmul(Rb[i][tid], Z64[i][tid]);
            pp_commit<PP_THREADS / 32>(p, buf, k, tb0, pass > 0, pass == 0);
        }
    }
}

// ---- layered v(tau) (:439-487): cumulative phase in fp64 turns (reduced mod 1), one sqrt + sincos per pair and tau
__global__ void __launch_bounds__(PP_THREADS, 2) phsh_layered_pair_kernel(const __grid_constant__ PhshParams p) {
    __shared__ float2 buf[PS_TB][PP_THREADS / 32];
    __shared__ double sv2[PS_TB], sth[PS_TB];
    const int tid = threadIdx.x;
    const int k = blockIdx.x;
    const int nh = max(1, p.nt / 2);
    const double kx = ps_kx(k, p.T, p.dx);
    // pairs with coss(tau = 0) <= thr2(0) are dead from the start (the mask is sticky, :484-485)
    int s_first = 0;
    {
        const double v0 = p.vmig[0], th0 = p.thr2[0];
        if (th0 < 1.0) {
            const double wmin = 0.5 * kx * fabs(v0) / sqrt(1.0 - th0);   // coss > th0  <=>  |w| > wmin
            s_first = (int)fmin((double)nh, floor(wmin * (double)p.nt * p.dt * 0.15915494309189535)) - 1;
            if (s_first < 0) s_first = 0;
        }
    }
    int pass = 0;
    for (int base = s_first; base < nh || pass == 0; base += PP_STATES, ++pass) {
        float2 A[PP_PER], B[PP_PER];
        double c[PP_PER], wt[PP_PER], ph[PP_PER];
        float plo[PP_PER];
#pragma unroll
        for (int i = 0; i < PP_PER; ++i) {
            const int s = base + tid + i * PP_THREADS;
            A[i] = make_float2(0.f, 0.f);
            B[i] = A[i];
            c[i] = 1e300;  // dead
            wt[i] = 0.0;
            ph[i] = 0.0;
            plo[i] = 0.f;
            if (s < nh) {
                const double w = ps_omega(s, p.nt, p.dt);
                const double h = 0.5 * kx / w;
                c[i] = h * h;                                // coss = 1 - c * v^2   (:460)
                wt[i] = w * p.dt * 0.15915494309189535;      // phase advance in turns per unit sqrt(coss), < 0.5
                pp_load_pair(p, s, k, A[i], B[i]);
            }
        }
        for (int tb0 = 0; tb0 < p.S; tb0 += PS_TB) {
            const int nb = min(PS_TB, p.S - tb0);
            if (tid < nb) {
                const double v = p.vmig[tb0 + tid];
                sv2[tid] = v * v;
                sth[tid] = p.thr2[tb0 + tid];
            }
            __syncthreads();
            for (int b = 0; b < nb; ++b) {
                const double v2 = sv2[b], th = sth[b];
                float2 part = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < PP_PER; ++i) {
                    const double coss = fma(-c[i], v2, 1.0);
                    if (coss <= th) {  // evanescent from here on (:484-485, sticky because FK itself is zeroed)
                        c[i] = 1e300;
                    } else {
                        // sqrt(coss) = s0 + e / (2 s0): fp32 rsqrt seed, fp64 residual; the small correction term is
                        // accumulated in fp32 next to the fp64 phase
                        const float cf = (float)coss;
                        float rs;
                        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(cf));
                        const float s0 = cf * rs;
                        const double s0d = (double)s0;
                        const float ef = (float)fma(-s0d, s0d, coss);
                        double phn = fma(wt[i], s0d, ph[i]);
                        if (phn > 0.5) phn -= 1.0;
                        ph[i] = phn;
                        plo[i] = fmaf((float)wt[i] * ef, 0.5f * rs, plo[i]);
                        float sn, cs;
                        __sincosf(((float)phn + plo[i]) * 6.283185307179586f, &sn, &cs);
                        part = pp_acc(part, cs, sn, A[i], B[i]);
                    }
                }
                pp_park<PP_THREADS / 32>(part, buf, b);
            }
            pp_commit<PP_THREADS / 32>(p, buf, k, tb0, pass > 0, false);
        }
    }
}

// The unpaired Nyquist-frequency bin of the layered case: FK[nt/2, k] cos(cumulative phase), one thread per kx.
__global__ void __launch_bounds__(128) phsh_layered_nyq_kernel(const __grid_constant__ PhshParams p) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p.K || p.nt < 2) return;
    const double w = ps_omega(p.nt / 2, p.nt, p.dt);
    const double h = 0.5 * ps_kx(k, p.T, p.dx) / w;
    const double c = h * h, wt = w * p.dt * 0.15915494309189535;
    const float2 f = p.FK[(size_t)(p.nt / 2) * p.K + k];
    double ph = 0.0;
    for (int tau = 0; tau < p.S; ++tau) {
        const double v = p.vmig[tau];
        const double coss = fma(-c, v * v, 1.0);
        if (coss <= p.thr2[tau]) break;
        ph = fma(wt, sqrt(coss), ph);
        ph -= rint(ph);
        const float cs = (float)cospi(2.0 * ph);
        float2 *dst = p.TK + (size_t)tau * p.K + k;
        float2 o = *dst;
        o.x = fmaf(f.x * cs, p.inv_s, o.x);
        o.y = fmaf(f.y * cs, p.inv_s, o.y);
        *dst = o;
    }
}

static int g_phsh_legacy = 0;  // testing hook: 0 auto (tensor-core constant velocity), 1 = the one-bin-per-state kernels, 3 = pair kernels only

struct PhshPlans {
    cufftHandle r2c = 0, c2c = 0, c2r = 0;
};
static std::map<std::tuple<int, int, int, int, long long>, PhshPlans> g_ps_plans;  // (device, S, T, nt, stream)
static std::mutex g_ps_mu;

static int ps_get_plans(int S, int T, int nt, cudaStream_t st, PhshPlans &out) {
    int dev = 0;
    IMPDAR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_ps_mu);
    auto key = std::make_tuple(dev, S, T, nt, (long long)(intptr_t)st);
    auto it = g_ps_plans.find(key);
    if (it != g_ps_plans.end()) {
        out = it->second;
        return IMPDAR_B200_OK;
    }
    const int K = T / 2 + 1;
    PhshPlans pl;
    {
        int n[1] = {T};
        IMPDAR_CUFFT(cufftPlanMany(&pl.r2c, 1, n, nullptr, 1, T, nullptr, 1, K, CUFFT_R2C, S));
    }
    {
        int n[1] = {nt}, emb[1] = {nt};
        IMPDAR_CUFFT(cufftPlanMany(&pl.c2c, 1, n, emb, K, 1, emb, K, 1, CUFFT_C2C, K));
    }
    {
        int n[1] = {T};
        IMPDAR_CUFFT(cufftPlanMany(&pl.c2r, 1, n, nullptr, 1, K, nullptr, 1, T, CUFFT_C2R, S));
    }
    g_ps_plans[key] = pl;
    out = pl;
    return IMPDAR_B200_OK;
}

static inline int ps_next_pow2(int S) {
    int nt = 1;
    while (nt < S) nt <<= 1;
    return nt;
}

}  // namespace impdar

using namespace impdar;

extern "C" {

size_t impdar_phsh_workspace_bytes(int S, int T) {
    const size_t nt = (size_t)ps_next_pow2(S);
    const size_t K = (size_t)(T / 2 + 1);
    const size_t fk = nt * K * sizeof(float2);
    const size_t tk = (size_t)S * K * sizeof(float2);
    const size_t real = (size_t)S * (size_t)T * sizeof(float);
    return fk + (tk > real ? tk : real) + phsh_tc_workspace_bytes((int)nt, (int)K) + 4096;
}

int impdar_phsh_f32(const float *data, float *out, int S, int T, double dt, double dx, double vel,
                    const double *vmig, const double *thr2, double htaper, double vtaper, void *workspace,
                    size_t ws_bytes, void *stream) {
    IMPDAR_CHECK_ARG(data && out, "phsh: null pointer");
    IMPDAR_CHECK_ARG(S >= 1 && T >= 1, "phsh: bad shape");
    IMPDAR_CHECK_ARG(dt > 0.0 && dx != 0.0, "phsh: dt must be positive and dx non-zero");
    IMPDAR_CHECK_ARG((vmig == nullptr) == (thr2 == nullptr), "phsh: vmig and thr2 go together");
    const size_t need = impdar_phsh_workspace_bytes(S, T);
    IMPDAR_CHECK_ARG(workspace && ws_bytes >= need, "phsh: workspace too small (%zu < %zu)", ws_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = ps_next_pow2(S);
    const int K = T / 2 + 1;
    const size_t fk_bytes = (size_t)nt * K * sizeof(float2);
    char *w = (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float2 *FK = (float2 *)w;
    char *w2 = w + ((fk_bytes + 255) & ~(size_t)255);
    float *tap = (float *)w2;    // tapered data, dead once the R2C has run
    float2 *TK = (float2 *)w2;   // ... so TK may overlay it

    PhshPlans pl;
    int rc = ps_get_plans(S, T, nt, st, pl);
    if (rc) return rc;
    IMPDAR_CUFFT(cufftSetStream(pl.r2c, st));
    IMPDAR_CUFFT(cufftSetStream(pl.c2c, st));
    IMPDAR_CUFFT(cufftSetStream(pl.c2r, st));

    rc = impdar_taper_f32(data, tap, S, T, 1, htaper, vtaper, 0, stream);
    if (rc) return rc;
    if (nt > S) IMPDAR_CUDA(cudaMemsetAsync(FK + (size_t)S * K, 0, (size_t)(nt - S) * K * sizeof(float2), st));
    IMPDAR_CUFFT(cufftExecR2C(pl.r2c, tap, (cufftComplex *)FK));                                   // x -> kx
    IMPDAR_CUFFT(cufftExecC2C(pl.c2c, (cufftComplex *)FK, (cufftComplex *)FK, CUFFT_FORWARD));     // t -> w
    count_launch(2);

    PhshParams p;
    p.FK = FK; p.TK = TK; p.nt = nt; p.K = K; p.S = S; p.T = T;
    p.dt = dt; p.dx = fabs(dx); p.vel = vel; p.vmig = vmig; p.thr2 = thr2;
    p.inv_s = (float)(1.0 / ((double)S * (double)T));  // /snum (:490-492) and numpy ifft's 1/tnum (:282)
    if ((g_phsh_legacy == 0 || g_phsh_legacy == 2) && vmig == nullptr && nt >= 64) {
        // constant velocity on the tensor cores (phaseshift_tc.cu)
        const size_t tk_bytes = (size_t)S * K * sizeof(float2), real_bytes = (size_t)S * T * sizeof(float);
        char *w3 = w2 + (((tk_bytes > real_bytes ? tk_bytes : real_bytes) + 255) & ~(size_t)255);
        rc = phsh_const_tc_launch(FK, TK, nt, K, S, T, dt, fabs(dx), vel, p.inv_s, w3, st);
        if (rc) return rc;
        IMPDAR_CUFFT(cufftExecC2R(pl.c2r, (cufftComplex *)TK, out));  // kx -> x
        count_launch(1);
        return IMPDAR_B200_OK;
    }
    if (g_phsh_legacy != 1) {
        if (vmig == nullptr) {
            ktimer_begin("phsh_const_pair_kernel", st);
            phsh_const_pair_kernel<<<K, PP_THREADS, 0, st>>>(p);
            ktimer_end(st);
            IMPDAR_LAUNCH_CHECK();
        } else {
            ktimer_begin("phsh_layered_pair_kernel", st);
            phsh_layered_pair_kernel<<<K, PP_THREADS, 0, st>>>(p);
            ktimer_end(st);
            IMPDAR_LAUNCH_CHECK();
            phsh_layered_nyq_kernel<<<(K + 127) / 128, 128, 0, st>>>(p);
            IMPDAR_LAUNCH_CHECK();
        }
        IMPDAR_CUFFT(cufftExecC2R(pl.c2r, (cufftComplex *)TK, out));  // kx -> x
        count_launch(1);
        return IMPDAR_B200_OK;
    }
    if (vmig == nullptr) {
        const size_t smem = 2 * (size_t)PSC_PER * PS_THREADS * sizeof(float2);
        IMPDAR_CUDA(cudaFuncSetAttribute(phsh_const_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ktimer_begin("phsh_const_kernel", st);
        phsh_const_kernel<<<(K + PSC_COLS - 1) / PSC_COLS, PS_THREADS, smem, st>>>(p);
        ktimer_end(st);
    } else {
        ktimer_begin("phsh_layered_kernel", st);
        phsh_layered_kernel<<<K, PS_THREADS, 0, st>>>(p);
        ktimer_end(st);
    }
    IMPDAR_LAUNCH_CHECK();
    IMPDAR_CUFFT(cufftExecC2R(pl.c2r, (cufftComplex *)TK, out));  // kx -> x (unnormalised)
    count_launch(1);
    return IMPDAR_B200_OK;
}

int impdar_phsh_set_legacy(int mode) {
    IMPDAR_CHECK_ARG(mode >= 0 && mode <= 3, "phsh_set_legacy: 0 automatic (constant velocity on the tensor cores - tcgen05, "
                                             "3xTF32 - layered velocity on the (+w, -w) pair kernel), 1 first-generation "
                                             "kernels, 2 = 0, 3 pair kernels for both");
    g_phsh_legacy = mode;
    return IMPDAR_B200_OK;
}

}  // extern "C"
