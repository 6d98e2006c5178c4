// Phase-shift migration with a laterally varying velocity v(x, z): split-step Fourier + explicit finite-difference
// correction ("Fourier finite-difference"), reference migrationlib/mig_python.py:428-432, 439-487 (phaseShift, 2-D
// vmig), :496-525 fourierFiniteDiff, :528-540 Sp_Matr.
//
// Per output time tau (sequential - the spectrum FK is carried from one tau to the next):
//   vbg = min_x vmig[tau], vfg = vmig[tau] - vbg, ufg = 1/vmig[tau] - 1/vbg                            (:452-454)
//   FK[w, k] *= exp(+i w dt Re sqrt(coss)),  coss = 1 - (vbg kx / 2w)^2                                (:460-464)
//   FFX[w, :] = ifft_k FK[w, :]                                                                        (:468)
//   FFX[w, x] *= exp(+i (2 ufg[x] w dt + vbg w dt))                        thin lens                   (:471-473)
//   tau > 0:  FFX[w] = L + c1 (A FFX[w]) + c2 (A FFX[w] - A L),  L = the row written last              (:476-478, :524)
//             c1 = dt alpha vfg^2 / (4 i w dx^2), c2 = -beta vfg^2 / (4 w^2 dx^2), alpha = 1/2, beta = 1/4
//   FK[w, :] = fft_x FFX[w, :];  FK[w, coss <= thr2[tau]] = 0;  TK[tau, k] = sum_w FK[w, k]            (:481-487)
// then out = Re ifft_k(TK / snum)                                                                      (:490-492, :282).
//
// Two properties of the reference shape this file:
//   * `FFX_last` is carried from one FREQUENCY to the next (:477-478), so the finite-difference update is one serial
//     chain of snum * nt steps.  All transforms and elementwise factors of a tau are batched over frequency (cuFFT Z2Z
//     over nt rows); only the O(tnum) stencil recurrence runs serially, in one persistent CTA per tau.
//   * Sp_Matr(tnum, -2, 1, 1) overwrites its main diagonal with zeros (the k3 = k4 = 0, nx = 0 setdiag calls, :534-535):
//     (A v)[0] = v[0], (A v)[i] = v[i-1] + v[i+1], (A v)[tnum-1] = sum(v)  (:537-540).
//
// The thin-lens phase reaches 1e8..1e9 radians (vbg * w * dt) and the chain compounds over snum * nt steps, so this
// branch runs in float64 end to end (complex128 spectra); it is latency-bound by construction, not bandwidth-bound.
#include <cufft.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace impdar {

#define IMPDAR_CUFFT(call)                                                                  \
    do {                                                                                    \
        cufftResult r__ = (call);                                                           \
        if (r__ != CUFFT_SUCCESS) {                                                         \
            impdar::set_error("%s:%d %s -> cufft error %d", __FILE__, __LINE__, #call, (int)r__); \
            return IMPDAR_B200_ECUFFT;                                                      \
        }                                                                                   \
    } while (0)

namespace ffd {

typedef double2 cd;

__device__ __forceinline__ cd cmul(cd a, cd b) {
    return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__device__ __forceinline__ cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }

__device__ __forceinline__ double taper_w(int i, int n, double len) {
    const int m = min(i, n - 1 - i);
    double w = (double)m / len;
    if (w > 1.0) w = 1.0;
    return w;
}

// tapered data (mig_python.py:253-258: dat.data *= H * V), zero-padded to nt rows, as complex128
__global__ void __launch_bounds__(256) taper_pad_kernel(const double *__restrict__ x, cd *__restrict__ z, int S, int T, int nt,
                                                        double htaper, double vtaper) {
    const long long n = (long long)nt * T;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(g / T), t = (int)(g % T);
        double v = 0.0;
        if (s < S) v = x[g] * __dmul_rn(taper_w(t, T, htaper), taper_w(s, S, vtaper));
        z[g] = make_double2(v, 0.0);
    }
}

// vbg[tau] = min_x vmig[tau, x]   (:452)
__global__ void __launch_bounds__(256) rowmin_kernel(const double *__restrict__ vmig, double *__restrict__ vbg, int T) {
    __shared__ double red[8];
    const double *row = vmig + (size_t)blockIdx.x * T;
    double m = INFINITY;
    bool nan = false;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const double v = row[t];
        nan |= (v != v);
        m = fmin(m, v);
    }
    if (nan) m = NAN;  // np.min propagates NaN
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double u = __shfl_xor_sync(0xffffffffu, m, o);
        m = (m != m || u != u) ? NAN : fmin(m, u);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = (m != m || red[i] != red[i]) ? NAN : fmin(m, red[i]);
        vbg[blockIdx.x] = m;
    }
}

// coss = 1 - (0.5 vbg kx / w)^2 with the reference's operation order (:460)
__device__ __forceinline__ double coss_of(double vbg, double kx, double w) {
    const double q = __ddiv_rn(__dmul_rn(__dmul_rn(0.5, vbg), kx), w);
    return __dsub_rn(1.0, __dmul_rn(q, q));
}

// FK[w, k] *= conj(cos(phase) + i sin(phase)), phase = (-w dt) Re sqrt(coss)   (:462-464)
__global__ void __launch_bounds__(256) shift_kernel(cd *__restrict__ FK, const double *__restrict__ ws, const double *__restrict__ kx,
                                                    const double *__restrict__ vbg_all, int tau, int nt, int T, double dt) {
    const double vbg = vbg_all[tau];
    const long long n = (long long)nt * T;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        const int iw = (int)(g / T), k = (int)(g % T);
        const double w = ws[iw];
        const double c = coss_of(vbg, kx[k], w);
        const double r = (c > 0.0) ? sqrt(c) : ((c == c) ? 0.0 : c);
        const double phase = __dmul_rn(__dmul_rn(-w, dt), r);
        double sn, cs;
        sincos(phase, &sn, &cs);
        FK[g] = cmul(FK[g], make_double2(cs, -sn));
    }
}

// FFX[w, x] = (1/T) * FFX[w, x] * exp(+i phase2), phase2 = 2 ufg w dt + 1 vbg w dt   (:468-473; the 1/T is numpy ifft's)
__global__ void __launch_bounds__(256) lens_kernel(cd *__restrict__ FFX, const double *__restrict__ ws, const double *__restrict__ vmig,
                                                   const double *__restrict__ vbg_all, int tau, int nt, int T, double dt) {
    const double vbg = vbg_all[tau];
    const double *vrow = vmig + (size_t)tau * T;
    const double invT = 1.0 / (double)T;
    const long long n = (long long)nt * T;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        const int iw = (int)(g / T), x = (int)(g % T);
        const double w = ws[iw];
        const double ufg = __dsub_rn(__ddiv_rn(1.0, vrow[x]), __ddiv_rn(1.0, vbg));
        const double p2 = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, ufg), w), dt), __dmul_rn(__dmul_rn(vbg, w), dt));
        double sn, cs;
        sincos(p2, &sn, &cs);
        const cd f = FFX[g];
        FFX[g] = cmul(make_double2(f.x * invT, f.y * invT), make_double2(cs, sn));
    }
}

// The serial chain over frequency (:476-478 with fourierFiniteDiff :517-524) for one tau > 0.  One CTA; F = thin-lens
// rows (nt, T), X = output rows, L = the row written last (Xl for iw = 0, X[iw - 1] afterwards).  On exit Xl = X[nt-1].
struct ChainSums {
    double fx, fy, lx, ly;
};

__global__ void __launch_bounds__(1024) chain_kernel(const cd *__restrict__ F, cd *X, cd *Xl,
                                                     const double *__restrict__ ws, const double *__restrict__ vmig,
                                                     const double *__restrict__ vbg_all, int tau, int nt, int T, double dt,
                                                     double dx) {
    __shared__ ChainSums red[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = (blockDim.x + 31) >> 5;
    const double vbg = vbg_all[tau];
    const double *vrow = vmig + (size_t)tau * T;
    const double alpha = 0.5, beta = 0.25;
    const double dx2 = dx * dx;
    for (int iw = 0; iw < nt; ++iw) {
        const double w = ws[iw];
        const cd *f = F + (size_t)iw * T;
        const cd *l = (iw == 0) ? Xl : X + (size_t)(iw - 1) * T;
        cd *o = X + (size_t)iw * T;
        // coeff1 = dt alpha vs^2 / (1j 4 w dx^2) = -i a1 vs^2,  coeff2 = -beta vs^2 / (4 w^2 dx^2) = -a2 vs^2
        const double a1 = dt * alpha / (4.0 * w * dx2), a2 = beta / (4.0 * w * w * dx2);
        ChainSums s = {0.0, 0.0, 0.0, 0.0};
        for (int x = tid; x < T; x += blockDim.x) {
            const cd fx = f[x], lx = l[x];
            s.fx += fx.x; s.fy += fx.y; s.lx += lx.x; s.ly += lx.y;
            if (x < T - 1) {
                cd af, al;
                if (x == 0) {
                    af = fx;
                    al = lx;
                } else {
                    af = cadd(f[x - 1], f[x + 1]);
                    al = cadd(l[x - 1], l[x + 1]);
                }
                const double vs = vrow[x] - vbg, g = vs * vs;
                const double c1 = -a1 * g, c2 = -a2 * g;  // coeff1 = i c1, coeff2 = c2
                const cd d = csub(af, al);
                o[x] = make_double2(lx.x - c1 * af.y + c2 * d.x, lx.y + c1 * af.x + c2 * d.y);
            }
        }
        // block sums of F and L for the all-ones last row of the stencil
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s.fx += __shfl_xor_sync(0xffffffffu, s.fx, off);
            s.fy += __shfl_xor_sync(0xffffffffu, s.fy, off);
            s.lx += __shfl_xor_sync(0xffffffffu, s.lx, off);
            s.ly += __shfl_xor_sync(0xffffffffu, s.ly, off);
        }
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (warp == 0) {
            ChainSums t = (lane < nwarp) ? red[lane] : ChainSums{0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                t.fx += __shfl_xor_sync(0xffffffffu, t.fx, off);
                t.fy += __shfl_xor_sync(0xffffffffu, t.fy, off);
                t.lx += __shfl_xor_sync(0xffffffffu, t.lx, off);
                t.ly += __shfl_xor_sync(0xffffffffu, t.ly, off);
            }
            if (lane == 0) {
                const int x = T - 1;
                const cd lx = l[x];
                cd af = make_double2(t.fx, t.fy), al = make_double2(t.lx, t.ly);
                if (T == 1) {  // row 0 and the last row coincide: A[-1, -1] = 1, A[-1, :-1] = 1 is written last (:539-540)
                    af = f[0];
                    al = lx;
                }
                const double vs = vrow[x] - vbg, g = vs * vs;
                const double c1 = -a1 * g, c2 = -a2 * g;
                const cd d = csub(af, al);
                o[x] = make_double2(lx.x - c1 * af.y + c2 * d.x, lx.y + c1 * af.x + c2 * d.y);
            }
        }
        __syncthreads();  // row iw is complete (global writes of this CTA are visible to it after the barrier)
    }
    const cd *last = X + (size_t)(nt - 1) * T;
    for (int x = tid; x < T; x += blockDim.x) Xl[x] = last[x];
}

// FK[w, coss <= thr2] = 0 (:484-485);  TK[tau, k] = sum_w FK[w, k] (:487).  blockDim = (32, 8): 32 wavenumbers x 8 slices of w.
__global__ void __launch_bounds__(256) mask_sum_kernel(cd *__restrict__ FK, cd *__restrict__ TK, const double *__restrict__ ws,
                                                       const double *__restrict__ kx, const double *__restrict__ vbg_all,
                                                       const double *__restrict__ thr2, int tau, int nt, int T) {
    __shared__ cd part[8][33];
    const int k = blockIdx.x * 32 + threadIdx.x;
    const double vbg = vbg_all[tau], thr = thr2[tau];
    cd acc = make_double2(0.0, 0.0);
    if (k < T) {
        const double kk = kx[k];
        for (int iw = threadIdx.y; iw < nt; iw += 8) {
            const size_t g = (size_t)iw * T + k;
            cd v = FK[g];
            if (coss_of(vbg, kk, ws[iw]) <= thr) {
                v = make_double2(0.0, 0.0);
                FK[g] = v;
            }
            acc = cadd(acc, v);
        }
    }
    part[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && k < T) {
        for (int j = 1; j < 8; ++j) acc = cadd(acc, part[j][threadIdx.x]);
        TK[(size_t)tau * T + k] = acc;
    }
}

// out = Re(ifft_k TK) / snum: the transform is unnormalised, so scale by 1 / (snum * tnum)   (:490-492, :282)
__global__ void __launch_bounds__(256) real_scale_kernel(const cd *__restrict__ Z, double *__restrict__ out, long long n, double sc) {
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x)
        out[g] = Z[g].x * sc;
}

struct Plans {
    cufftHandle fft2 = 0, rows_nt = 0, rows_s = 0;
};
static std::map<std::tuple<int, int, int, int, long long>, Plans> g_plans;  // (device, S, T, nt, stream)
static std::mutex g_mu;

static int get_plans(int S, int T, int nt, cudaStream_t st, Plans &out) {
    int dev = 0;
    IMPDAR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(dev, S, T, nt, (long long)(intptr_t)st);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) {
        out = it->second;
        return IMPDAR_B200_OK;
    }
    Plans pl;
    IMPDAR_CUFFT(cufftPlan2d(&pl.fft2, nt, T, CUFFT_Z2Z));
    int n[1] = {T};
    IMPDAR_CUFFT(cufftPlanMany(&pl.rows_nt, 1, n, nullptr, 1, T, nullptr, 1, T, CUFFT_Z2Z, nt));
    IMPDAR_CUFFT(cufftPlanMany(&pl.rows_s, 1, n, nullptr, 1, T, nullptr, 1, T, CUFFT_Z2Z, S));
    IMPDAR_CUFFT(cufftSetStream(pl.fft2, st));
    IMPDAR_CUFFT(cufftSetStream(pl.rows_nt, st));
    IMPDAR_CUFFT(cufftSetStream(pl.rows_s, st));
    g_plans[key] = pl;
    out = pl;
    return IMPDAR_B200_OK;
}

static inline int next_pow2(int S) {
    int nt = 1;
    while (nt < S) nt <<= 1;
    return nt;
}
static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

static inline unsigned grid_for(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace ffd
}  // namespace impdar

using namespace impdar;

extern "C" {

size_t impdar_phsh_ffd_workspace_bytes(int S, int T) {
    const size_t nt = (size_t)ffd::next_pow2(S);
    const size_t plane = ffd::al256(nt * (size_t)T * sizeof(double2));
    const size_t tk = ffd::al256((size_t)S * (size_t)T * sizeof(double2));
    return 3 * plane + tk + ffd::al256((size_t)T * sizeof(double2)) + ffd::al256((size_t)T * sizeof(double)) +
           ffd::al256(nt * sizeof(double)) + ffd::al256((size_t)S * sizeof(double)) + 512;
}

int impdar_phsh_ffd_f64(const double *data, double *out, int S, int T, double dt, double dx, double dx_fd, const double *vmig,
                        const double *thr2, double htaper, double vtaper, void *workspace, size_t ws_bytes, void *stream) {
    using namespace ffd;
    IMPDAR_CHECK_ARG(data && out && vmig && thr2, "phsh_ffd: null pointer");
    IMPDAR_CHECK_ARG(S >= 1 && T >= 1, "phsh_ffd: bad shape");
    IMPDAR_CHECK_ARG(dt > 0.0 && dx != 0.0 && dx_fd != 0.0, "phsh_ffd: dt must be positive and dx, dx_fd non-zero");
    const size_t need = impdar_phsh_ffd_workspace_bytes(S, T);
    IMPDAR_CHECK_ARG(workspace && ws_bytes >= need, "phsh_ffd: workspace too small (%zu < %zu)", ws_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = next_pow2(S);
    const size_t plane = al256((size_t)nt * T * sizeof(cd));
    char *w = (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    cd *FK = (cd *)w; w += plane;
    cd *FFX = (cd *)w; w += plane;
    cd *X = (cd *)w; w += plane;
    cd *TK = (cd *)w; w += al256((size_t)S * T * sizeof(cd));
    cd *Xl = (cd *)w; w += al256((size_t)T * sizeof(cd));
    double *kx_d = (double *)w; w += al256((size_t)T * sizeof(double));
    double *ws_d = (double *)w; w += al256((size_t)nt * sizeof(double));
    double *vbg_d = (double *)w;

    // kx = 2 pi fftfreq(tnum, dx), ws = 2 pi fftfreq(nt, dt) with w == 0 -> 1e-10 / dt   (:263-266, :447-448)
    std::vector<double> kx_h(T), ws_h(nt);
    auto fftfreq = [](std::vector<double> &v, int n, double d) {
        const double val = 1.0 / ((double)n * d);
        const int npos = (n - 1) / 2 + 1;
        for (int i = 0; i < n; ++i) v[i] = (2.0 * M_PI) * ((double)(i < npos ? i : i - n) * val);
    };
    fftfreq(kx_h, T, dx);
    fftfreq(ws_h, nt, dt);
    for (int i = 0; i < nt; ++i)
        if (ws_h[i] == 0.0) ws_h[i] = 1.0e-10 / dt;
    IMPDAR_CUDA(cudaMemcpyAsync(kx_d, kx_h.data(), (size_t)T * sizeof(double), cudaMemcpyHostToDevice, st));
    IMPDAR_CUDA(cudaMemcpyAsync(ws_d, ws_h.data(), (size_t)nt * sizeof(double), cudaMemcpyHostToDevice, st));
    IMPDAR_CUDA(cudaStreamSynchronize(st));  // the host vectors die with this frame

    Plans pl;
    int rc = get_plans(S, T, nt, st, pl);
    if (rc) return rc;

    const long long n_plane = (long long)nt * T;
    taper_pad_kernel<<<grid_for(n_plane), 256, 0, st>>>(data, FK, S, T, nt, htaper, vtaper);
    IMPDAR_LAUNCH_CHECK();
    rowmin_kernel<<<S, 256, 0, st>>>(vmig, vbg_d, T);
    IMPDAR_LAUNCH_CHECK();
    IMPDAR_CUFFT(cufftExecZ2Z(pl.fft2, (cufftDoubleComplex *)FK, (cufftDoubleComplex *)FK, CUFFT_FORWARD));   // :270
    count_launch(1);
    IMPDAR_CUDA(cudaMemsetAsync(Xl, 0, (size_t)T * sizeof(cd), st));

    int chain_threads = ((T + 31) / 32) * 32;
    if (chain_threads > 1024) chain_threads = 1024;
    const dim3 ms_block(32, 8), ms_grid((T + 31) / 32);
    for (int tau = 0; tau < S; ++tau) {
        shift_kernel<<<grid_for(n_plane), 256, 0, st>>>(FK, ws_d, kx_d, vbg_d, tau, nt, T, dt);
        IMPDAR_LAUNCH_CHECK();
        IMPDAR_CUFFT(cufftExecZ2Z(pl.rows_nt, (cufftDoubleComplex *)FK, (cufftDoubleComplex *)FFX, CUFFT_INVERSE));
        lens_kernel<<<grid_for(n_plane), 256, 0, st>>>(FFX, ws_d, vmig, vbg_d, tau, nt, T, dt);
        IMPDAR_LAUNCH_CHECK();
        const cd *src = FFX;
        if (tau > 0) {
            ktimer_begin("chain_kernel", st);
            chain_kernel<<<1, chain_threads, 0, st>>>(FFX, X, Xl, ws_d, vmig, vbg_d, tau, nt, T, dt, fabs(dx_fd));
            ktimer_end(st);
            IMPDAR_LAUNCH_CHECK();
            src = X;
        } else {
            IMPDAR_CUDA(cudaMemcpyAsync(Xl, FFX + (size_t)(nt - 1) * T, (size_t)T * sizeof(cd), cudaMemcpyDeviceToDevice, st));
        }
        IMPDAR_CUFFT(cufftExecZ2Z(pl.rows_nt, (cufftDoubleComplex *)src, (cufftDoubleComplex *)FK, CUFFT_FORWARD));
        count_launch(2);
        mask_sum_kernel<<<ms_grid, ms_block, 0, st>>>(FK, TK, ws_d, kx_d, vbg_d, thr2, tau, nt, T);
        IMPDAR_LAUNCH_CHECK();
    }
    IMPDAR_CUFFT(cufftExecZ2Z(pl.rows_s, (cufftDoubleComplex *)TK, (cufftDoubleComplex *)TK, CUFFT_INVERSE));
    count_launch(1);
    real_scale_kernel<<<grid_for((long long)S * T), 256, 0, st>>>(TK, out, (long long)S * T, 1.0 / ((double)S * (double)T));
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

}  // extern "C"
