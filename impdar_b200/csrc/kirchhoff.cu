// Kirchhoff diffraction summation (reference: migrationlib/mig_python.py:35-123, C prototype
// migrationlib/mig_cython.h:11).
//
// out[ti, xi] = 1/(2 pi) * nansum_x( gradD[idx(ti,xi,x), x] * cos(theta)/v  [+ data[idx, x] * cos(theta)/r^2] )
//   r = sqrt((dist[x]-dist[xi])^2 + zs[ti]^2), t = 2 r / v, idx = argmin_k |tt[k] - t| (first minimum),
//   terms with t > max(tt) are zeroed (mig_python.py:52), NaN terms are skipped (nansum, :53).
//
// Design (general geometry kernel):
//   * pre-pass: d/dt (np.gradient's stencil, fp64) fused with a transpose to trace-major gradT[x][k],
//     zero-padded past the last sample, so one warp's gather (32 consecutive output samples, one input
//     trace) touches one or two 128-byte lines and is bank/sector conflict free.
//   * main kernel: a warp owns 32 consecutive output samples of one output trace; a CTA owns 8 adjacent
//     output traces, so the 8 warps gather from the same lines.  Per (output sample, input trace) pair the
//     travel time is evaluated in sample units in fp32: q = U^2 + C^2, s = rsqrt(q), f = q*s - U0, the
//     weight U*s/v falls out of the same rsqrt.  The nearest-sample pick round(f) is only trusted when f
//     is further than `delta` from a rounding boundary; the rare ambiguous pairs are queued per warp and
//     re-evaluated with the reference's exact float64 operation sequence (sqrt, div, |tt[k]-t| compares),
//     so the picked sample is the reference's for every pair.
//   * the aperture cut 2r/v > max(tt) is resolved exactly once per output sample (binary search with the
//     float64 predicate over the ascending trace positions), so the inner loop only tests a trace range.
#include "common.cuh"

namespace impdar {

struct KirchParams {
    const float *gradT;   // (T, SP)
    const float *dataT;   // (T, SP) or null
    float *out;           // (S, ldo)
    const double *dist;   // T  [m]
    const double *tt;     // S  [s]
    const double *zs;     // S  vel*tt/2
    const double *zs2;    // S  zs^2
    const int *flags;     // [0] = non-finite input seen
    unsigned long long *stats;  // [0] pairs, [1] exact pairs
    int S, T, SP, ldo, x_begin, x_end;
    double vel, tmax, tt0, inv_dt, cs;  // cs = 2/(vel*dt_eff)
    float neg_u0, thr, wfar, wnear;     // thr = 0.5 - delta ; wfar = 1/(2 pi vel) ; wnear = 4/(vel dt)^2/(2 pi)
    int monotone;                       // dist ascending -> per-sample aperture by binary search
};

#define KIRCH_MAGIC 12582912.0f  // 1.5 * 2^23
#define KIRCH_MAGIC_I 0x4B400000

__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// The reference's own float64 sequence for one pair: returns t = 2*sqrt(d^2 + zs2)/vel without contraction.
__device__ __forceinline__ double exact_time(double d, double zs2, double vel, double &rs) {
    rs = __dsqrt_rn(__dadd_rn(__dmul_rn(d, d), zs2));
    return __ddiv_rn(__dmul_rn(2.0, rs), vel);
}

// argmin_k |tt[k] - t| with first-minimum ties, tt ascending.
__device__ __forceinline__ int exact_pick(const double *__restrict__ tt, int S, double t, double tt0,
                                          double inv_dt) {
    double g = (t - tt0) * inv_dt;
    int k = (g < 0.0) ? 0 : ((g > (double)(S - 1)) ? S - 1 : (int)rint(g));
    double dk = fabs(__dsub_rn(tt[k], t));
    while (k > 0) {
        double dm = fabs(__dsub_rn(tt[k - 1], t));
        if (dm <= dk) { dk = dm; --k; } else break;
    }
    while (k < S - 1) {
        double dp = fabs(__dsub_rn(tt[k + 1], t));
        if (dp < dk) { dk = dp; ++k; } else break;
    }
    return k;
}

template <bool NEAR>
__device__ __forceinline__ float exact_term(const KirchParams &p, int ti, int x, double dxi) {
    const double d = __dsub_rn(p.dist[x], dxi);
    double rs;
    const double t = exact_time(d, p.zs2[ti], p.vel, rs);
    if (t > p.tmax) return 0.f;
    const int k = exact_pick(p.tt, p.S, t, p.tt0, p.inv_dt);
    const double costheta = p.zs[ti] / rs;
    const double g = (double)p.gradT[(size_t)x * p.SP + k];
    double term = g * costheta / p.vel;
    if (term != term) term = 0.0;
    if (NEAR) {
        const double dv = (double)p.dataT[(size_t)x * p.SP + k];
        double tn = dv * costheta / (rs * rs);
        if (tn == tn) term += tn;
    }
    return (float)(term * 0.15915494309189535);  // 1/(2 pi)
}

#define KQ_CAP 192

template <bool NEAR, bool STATS>
__global__ void __launch_bounds__(256) kirch_general_kernel(const __grid_constant__ KirchParams p) {
    __shared__ float sC[8][32];
    __shared__ int sQ[8][KQ_CAP];
    __shared__ float sT[8][32];
    __shared__ int sO[8][32];
    __shared__ float sOut[32][9];

    const int lane = threadIdx.x, warp = threadIdx.y;
    const int ti0 = blockIdx.x * 32;
    const int ti = ti0 + lane;
    const int xi = p.x_begin + blockIdx.y * 8 + warp;
    const bool warp_active = xi < p.x_end;
    const bool active = warp_active && ti < p.S;
    const bool hasnan = p.flags[0] != 0;

    float acc0 = 0.f, acc1 = 0.f, accx = 0.f;
    unsigned npairs = 0, nexact = 0;

    if (warp_active) {
        const double dxi = p.dist[xi];
        const int tic = active ? ti : p.S - 1;
        const double zs2i = p.zs2[tic];
        const float U = (float)(p.tt[tic] * p.inv_dt);
        float u2 = (float)((p.tt[tic] * p.inv_dt) * (p.tt[tic] * p.inv_dt));
        u2 = fmaxf(u2, 1e-30f);  // q == 0 only at the apex of a zero-depth sample: weight is 0 there

        // exact per-sample aperture [xlo, xhi]: pairs with 2r/v > max(tt) contribute exactly 0
        int xlo = xi + 1, xhi = xi;
        if (active) {
            double rs;
            if (!(exact_time(0.0, zs2i, p.vel, rs) > p.tmax)) {
                if (p.monotone) {
                    int a = 0, b = xi;  // smallest x in [0, xi] inside
                    while (a < b) {
                        int m = (a + b) >> 1;
                        if (exact_time(__dsub_rn(p.dist[m], dxi), zs2i, p.vel, rs) > p.tmax) a = m + 1; else b = m;
                    }
                    xlo = a;
                    a = xi; b = p.T - 1;  // largest x in [xi, T-1] inside
                    while (a < b) {
                        int m = (a + b + 1) >> 1;
                        if (exact_time(__dsub_rn(p.dist[m], dxi), zs2i, p.vel, rs) > p.tmax) b = m - 1; else a = m;
                    }
                    xhi = a;
                } else {
                    xlo = 0;
                    xhi = p.T - 1;
                }
            }
        }
        int xlo_w = xlo, xhi_w = xhi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            xlo_w = min(xlo_w, __shfl_xor_sync(0xffffffffu, xlo_w, o));
            xhi_w = max(xhi_w, __shfl_xor_sync(0xffffffffu, xhi_w, o));
        }
        const unsigned span = (unsigned)(xhi - xlo);  // xhi < xlo -> huge unsigned? no: xlo = xi+1, xhi = xi -> 0xffffffff
        const bool lane_empty = xhi < xlo;
        const float thr = p.monotone ? p.thr : -1.f;  // non-monotone positions: every pair takes the exact path
        int qn = 0;

        auto flush = [&](int count) {
            for (int base = 0; base < count; base += 32) {
                float term = 0.f;
                int owner = -1;
                if (base + lane < count) {
                    const int e = sQ[warp][base + lane];
                    owner = e & 31;
                    term = exact_term<NEAR>(p, ti0 + owner, e >> 5, dxi);
                }
                sT[warp][lane] = term;
                sO[warp][lane] = owner;
                __syncwarp();
                const int n = min(32, count - base);
                for (int j = 0; j < n; ++j)
                    if (sO[warp][j] == lane) accx += sT[warp][j];
                __syncwarp();
            }
        };

        for (int xb = xlo_w; xb <= xhi_w; xb += 32) {
            {
                const int xx = xb + lane;
                float c2v = 0.f;
                if (xx <= xhi_w) {
                    const double C = (p.dist[xx] - dxi) * p.cs;
                    c2v = (float)(C * C);
                }
                __syncwarp();
                sC[warp][lane] = c2v;
                __syncwarp();
            }
            const int n = min(32, xhi_w - xb + 1);
            const float *gp = p.gradT + (size_t)xb * p.SP;
            const float *dp = NEAR ? p.dataT + (size_t)xb * p.SP : nullptr;
            for (int j0 = 0; j0 < n; j0 += 4) {
                unsigned ambmask = 0;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = j0 + jj;
                    if (j < n) {
                        const float c2 = sC[warp][j];
                        const float q = u2 + c2;
                        const float s = rsqrt_approx(q);
                        const float f = fmaf(q, s, p.neg_u0);
                        const float r = f + KIRCH_MAGIC;
                        const int k = (int)min((unsigned)(__float_as_int(r) - KIRCH_MAGIC_I), (unsigned)(p.SP - 1));
                        const float e = f - (r - KIRCH_MAGIC);
                        const bool inr = !lane_empty && ((unsigned)(xb + j - xlo) <= span);
                        const bool amb = fabsf(e) > thr;
                        const float g = __ldg(gp + (size_t)j * p.SP + k);
                        const float w = U * s;
                        if (inr && !amb) {
                            float term = g * w * p.wfar;
                            if (NEAR) {
                                const float dv = __ldg(dp + (size_t)j * p.SP + k);
                                float tn = dv * ((w * p.wnear) * (s * s));
                                if (hasnan && tn != tn) tn = 0.f;
                                if (hasnan && term != term) term = 0.f;
                                term += tn;
                            } else if (hasnan && term != term) {
                                term = 0.f;
                            }
                            if (jj & 1) acc1 += term; else acc0 += term;
                        }
                        if (STATS && inr) ++npairs;
                        if (inr && amb) ambmask |= 1u << jj;
                    }
                }
                if (__any_sync(0xffffffffu, ambmask != 0)) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const bool a = (ambmask >> jj) & 1u;
                        const unsigned m = __ballot_sync(0xffffffffu, a);
                        if (a) {
                            sQ[warp][qn + __popc(m & ((1u << lane) - 1u))] = ((xb + j0 + jj) << 5) | lane;
                            ++nexact;
                        }
                        qn += __popc(m);
                    }
                    __syncwarp();
                    if (qn > 32) {
                        flush(qn);
                        qn = 0;
                    }
                }
            }
        }
        __syncwarp();
        if (qn > 0) flush(qn);
    }

    // transpose the 32x8 tile so that rows of `out` are written in 32-byte runs
    sOut[lane][warp] = acc0 + acc1 + accx;
    __syncthreads();
    {
        const int tid = warp * 32 + lane;
        const int r = tid >> 3, c = tid & 7;
        const int to = ti0 + r, xo = p.x_begin + blockIdx.y * 8 + c;
        if (to < p.S && xo < p.x_end) p.out[(size_t)to * p.ldo + (xo - p.x_begin)] = sOut[r][c];
    }
    if (STATS) {
        unsigned long long np64 = npairs, ne64 = nexact;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            np64 += __shfl_xor_sync(0xffffffffu, np64, o);
            ne64 += __shfl_xor_sync(0xffffffffu, ne64, o);
        }
        if (lane == 0 && np64) {
            atomicAdd(&p.stats[0], np64);
            if (ne64) atomicAdd(&p.stats[1], ne64);
        }
    }
}

// d/dt with np.gradient's stencil (coefficients a,b,c per row; b == 0 in numpy's uniform branch, which
// never touches f[s]) fused with the transpose to trace-major and a non-finite scan.
__global__ void __launch_bounds__(256) grad_transpose_kernel(const float *__restrict__ data,
                                                             float *__restrict__ gradT,
                                                             float *__restrict__ dataT, int S, int T, int SP,
                                                             const double *__restrict__ coef,
                                                             int *__restrict__ flags) {
    __shared__ float tg[32][33];
    __shared__ float td[32][33];
    const int s0 = blockIdx.y * 32, x0 = blockIdx.x * 32;
    bool bad = false;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int s = s0 + r, x = x0 + threadIdx.x;
        float g = 0.f, dv = 0.f;
        if (s < S && x < T) {
            const double a = coef[s], b = coef[S + s], c = coef[2 * S + s];
            dv = data[(size_t)s * T + x];
            double acc = 0.0;
            if (s > 0 && a != 0.0) acc += a * (double)data[(size_t)(s - 1) * T + x];
            if (b != 0.0) acc += b * (double)dv;
            if (s < S - 1 && c != 0.0) acc += c * (double)data[(size_t)(s + 1) * T + x];
            g = (float)acc;
            if (!isfinite(g) || !isfinite(dv)) bad = true;
        }
        tg[r][threadIdx.x] = g;
        td[r][threadIdx.x] = dv;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int x = x0 + r, s = s0 + threadIdx.x;
        if (x < T && s < S) {
            gradT[(size_t)x * SP + s] = tg[threadIdx.x][r];
            if (dataT) dataT[(size_t)x * SP + s] = td[threadIdx.x][r];
        }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0 && threadIdx.y == 0) atomicOr(flags, 1);
}

__global__ void kirch_prep_vectors_kernel(const double *__restrict__ tt, double *__restrict__ zs,
                                          double *__restrict__ zs2, int S, double vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) {
        const double z = __ddiv_rn(__dmul_rn(vel, tt[i]), 2.0);  // zs = vel * tt_sec / 2.0   (:101)
        zs[i] = z;
        zs2[i] = __dmul_rn(z, z);                                // zs2 = zs**2.              (:102)
    }
}

static inline int kirch_sp(int S) { return ((S + 64 + 31) / 32) * 32; }

static unsigned long long *g_last_stats = nullptr;
static cudaStream_t g_last_stats_stream = nullptr;
static int g_stats_enabled = 0;

}  // namespace impdar

using namespace impdar;

extern "C" {

size_t impdar_kirchhoff_workspace_bytes(int S, int T, int nearfield) {
    const size_t sp = (size_t)kirch_sp(S);
    size_t b = (size_t)T * sp * sizeof(float) * (nearfield ? 2 : 1);
    b += 6 * (size_t)S * sizeof(double);  // zs, zs2, tt, grad coefficients
    b += (size_t)T * sizeof(double);      // dist
    b += 256;                              // flags + stats
    return b + 1024;
}

int impdar_kirchhoff_f32(const float *data, float *out, int S, int T, const double *dist_m, const double *tt_s,
                         const double *grad_coef, double vel, int nearfield, int x_begin, int x_end,
                         void *workspace, size_t ws_bytes, void *stream) {
    IMPDAR_CHECK_ARG(data && out && dist_m && tt_s && grad_coef, "kirchhoff: null pointer");
    IMPDAR_CHECK_ARG(S >= 2 && T >= 1, "kirchhoff: need snum >= 2, tnum >= 1");
    IMPDAR_CHECK_ARG(0 <= x_begin && x_begin < x_end && x_end <= T, "kirchhoff: bad output range [%d, %d)",
                     x_begin, x_end);
    IMPDAR_CHECK_ARG(vel > 0.0, "kirchhoff: vel must be positive");
    IMPDAR_CHECK_ARG(T < (1 << 26), "kirchhoff: tnum too large");
    const size_t need = impdar_kirchhoff_workspace_bytes(S, T, nearfield);
    IMPDAR_CHECK_ARG(workspace && ws_bytes >= need, "kirchhoff: workspace too small (%zu < %zu)", ws_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;

    const double *h_tt = tt_s, *h_dist = dist_m;  // HOST vectors (small); device copies live in the workspace
    double tmax = h_tt[0];
    bool ascending = true;
    for (int i = 0; i < S; ++i) {
        if (h_tt[i] > tmax) tmax = h_tt[i];
        if (i > 0 && !(h_tt[i] > h_tt[i - 1])) ascending = false;
    }
    IMPDAR_CHECK_ARG(ascending, "kirchhoff: travel_time must be strictly ascending");
    const double tt0 = h_tt[0];
    const double dt_eff = (h_tt[S - 1] - h_tt[0]) / (double)(S - 1);
    double maxdev = 0.0;
    for (int i = 0; i < S; ++i) {
        const double dev = fabs(h_tt[i] - (tt0 + i * dt_eff));
        if (dev > maxdev) maxdev = dev;
    }
    int monotone = 1;
    for (int i = 1; i < T; ++i)
        if (!(h_dist[i] >= h_dist[i - 1])) { monotone = 0; break; }
    // fp32 error budget of f = sqrt(U^2 + C^2) - U0 (sample units): ~4 ulp relative on a value <= S + |U0|,
    // plus the deviation of tt from a uniform grid.
    const double u0 = tt0 / dt_eff;
    double delta = 6e-7 * ((double)S + fabs(u0) + 64.0) + 2.0 * maxdev / dt_eff + 1e-6;
    float thr = (float)(0.5 - delta);
    if (delta > 0.2) thr = -1.f;  // irregular sampling: everything through the exact path

    // carve workspace
    const int SP = kirch_sp(S);
    char *w = (char *)workspace;
    w = (char *)(((uintptr_t)w + 255) & ~(uintptr_t)255);
    float *gradT = (float *)w;
    w += (size_t)T * SP * sizeof(float);
    float *dataT = nullptr;
    if (nearfield) {
        dataT = (float *)w;
        w += (size_t)T * SP * sizeof(float);
    }
    double *zs = (double *)w;
    w += (size_t)S * sizeof(double);
    double *zs2 = (double *)w;
    w += (size_t)S * sizeof(double);
    double *d_tt = (double *)w;
    w += (size_t)S * sizeof(double);
    double *d_coef = (double *)w;
    w += 3 * (size_t)S * sizeof(double);
    double *d_dist = (double *)w;
    w += (size_t)T * sizeof(double);
    IMPDAR_CUDA(cudaMemcpyAsync(d_tt, tt_s, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, st));
    IMPDAR_CUDA(cudaMemcpyAsync(d_coef, grad_coef, 3 * (size_t)S * sizeof(double), cudaMemcpyHostToDevice, st));
    IMPDAR_CUDA(cudaMemcpyAsync(d_dist, dist_m, (size_t)T * sizeof(double), cudaMemcpyHostToDevice, st));
    int *flags = (int *)w;
    unsigned long long *stats = (unsigned long long *)(w + 64);

    IMPDAR_CUDA(cudaMemsetAsync(gradT, 0, (size_t)T * SP * sizeof(float) * (nearfield ? 2 : 1), st));
    IMPDAR_CUDA(cudaMemsetAsync(flags, 0, 128, st));
    kirch_prep_vectors_kernel<<<(S + 255) / 256, 256, 0, st>>>(d_tt, zs, zs2, S, vel);
    IMPDAR_LAUNCH_CHECK();
    {
        dim3 grid((T + 31) / 32, (S + 31) / 32), block(32, 8);
        grad_transpose_kernel<<<grid, block, 0, st>>>(data, gradT, dataT, S, T, SP, d_coef, flags);
        IMPDAR_LAUNCH_CHECK();
    }
    KirchParams p;
    p.gradT = gradT; p.dataT = dataT; p.out = out; p.dist = d_dist; p.tt = d_tt; p.zs = zs; p.zs2 = zs2;
    p.flags = flags; p.stats = stats;
    p.S = S; p.T = T; p.SP = SP; p.ldo = x_end - x_begin; p.x_begin = x_begin; p.x_end = x_end;
    p.vel = vel; p.tmax = tmax; p.tt0 = tt0; p.inv_dt = 1.0 / dt_eff; p.cs = 2.0 / (vel * dt_eff);
    p.neg_u0 = (float)(-u0); p.thr = thr;
    p.wfar = (float)(1.0 / (2.0 * 3.14159265358979323846 * vel));
    p.wnear = (float)(4.0 / ((vel * dt_eff) * (vel * dt_eff)) / (2.0 * 3.14159265358979323846));
    p.monotone = monotone;
    dim3 grid((S + 31) / 32, (x_end - x_begin + 7) / 8), block(32, 8);
    if (g_stats_enabled) {
        if (nearfield) kirch_general_kernel<true, true><<<grid, block, 0, st>>>(p);
        else kirch_general_kernel<false, true><<<grid, block, 0, st>>>(p);
    } else {
        if (nearfield) kirch_general_kernel<true, false><<<grid, block, 0, st>>>(p);
        else kirch_general_kernel<false, false><<<grid, block, 0, st>>>(p);
    }
    IMPDAR_LAUNCH_CHECK();
    g_last_stats = stats;
    g_last_stats_stream = st;
    return IMPDAR_B200_OK;
}

int impdar_kirchhoff_enable_stats(int on) {
    g_stats_enabled = on ? 1 : 0;
    return IMPDAR_B200_OK;
}

int impdar_kirchhoff_last_stats(unsigned long long *pairs, unsigned long long *exact_pairs) {
    IMPDAR_CHECK_ARG(g_last_stats, "kirchhoff_last_stats: no previous call");
    unsigned long long h[2];
    IMPDAR_CUDA(cudaMemcpyAsync(h, g_last_stats, sizeof(h), cudaMemcpyDeviceToHost, g_last_stats_stream));
    IMPDAR_CUDA(cudaStreamSynchronize(g_last_stats_stream));
    if (pairs) *pairs = h[0];
    if (exact_pairs) *exact_pairs = h[1];
    return IMPDAR_B200_OK;
}

}  // extern "C"
