// Filtering kernels: taper, horizontalfilt, adaptivehfilt, filtfilt (vertical_band_pass), FIR lfilter.
// Reference semantics: RadarData/_RadarDataFiltering.py (line numbers in include/impdar_b200.h).
// All of these are HBM-bound: traces are the contiguous axis, so every kernel maps lanes to traces
// (coalesced 128-bit accesses) and keeps per-row / per-trace state on chip.
#include <stdlib.h>

#include "common.cuh"
#include "iir.cuh"

namespace impdar {

// --------------------------------------------------------------------------------------------- taper
__device__ __forceinline__ double taper_weight(int i, int n, double len) {
    int m = min(i, n - 1 - i);
    double w = (double)m / len;  // 0/0 -> NaN, m/0 -> inf -> clipped to 1, like numpy
    if (w > 1.0) w = 1.0;
    return w;
}

__global__ void __launch_bounds__(256) taper_kernel(const float *__restrict__ x, float *__restrict__ y, int S,
                                                    int T, long long rows, double htaper, double vtaper,
                                                    int trunc_int) {
    // one row per blockIdx.y-stride, columns vectorised by 4 when possible
    const bool vec = ((T & 3) == 0) && ((((uintptr_t)x | (uintptr_t)y) & 15) == 0);
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int s = (int)(row % S);
        const double v = taper_weight(s, S, vtaper);
        const float *xr = x + row * (long long)T;
        float *yr = y + row * (long long)T;
        if (vec) {
            for (int c = threadIdx.x * 4; c < T; c += blockDim.x * 4) {
                float4 a = ld_stream4(xr + c);
                float r[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double p = (double)r[j] * taper_weight(c + j, T, htaper) * v;
                    if (trunc_int) p = trunc(p);
                    r[j] = (float)p;
                }
                st_stream4(yr + c, make_float4(r[0], r[1], r[2], r[3]));
            }
        } else {
            for (int c = threadIdx.x; c < T; c += blockDim.x) {
                double p = (double)xr[c] * taper_weight(c, T, htaper) * v;
                if (trunc_int) p = trunc(p);
                yr[c] = (float)p;
            }
        }
    }
}

// float64 radargrams (the time-wavenumber stub multiplies a float64 array in place, mig_python.py:330-335): one rounding
// per product in the reference's own order, data * (H * V)
__global__ void __launch_bounds__(256) taper_f64_kernel(const double *__restrict__ x, double *__restrict__ y, int S, int T,
                                                        long long rows, double htaper, double vtaper) {
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const double v = taper_weight((int)(row % S), S, vtaper);
        const double *xr = x + row * (long long)T;
        double *yr = y + row * (long long)T;
        for (int c = threadIdx.x; c < T; c += blockDim.x) yr[c] = __dmul_rn(xr[c], __dmul_rn(taper_weight(c, T, htaper), v));
    }
}

// -------------------------------------------------------------------------------------------- hfilt
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
    typedef float4 type;
    static constexpr int N = 4;
};
template <>
struct Vec4<double> {
    typedef double2 type;
    static constexpr int N = 2;
};

__device__ __forceinline__ double block_sum(double v, double *red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();  // protect red[] from the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double t = (lane < nw) ? red[lane] : 0.0;
    t = warp_sum(t);
    return t;
}

template <typename T>
__global__ void __launch_bounds__(256) hfilt_kernel(const T *__restrict__ x, T *__restrict__ y, int S, int Tn,
                                                    long long rows, int htr1, int htrn,
                                                    const double *__restrict__ taper, int trunc_avg) {
    __shared__ double red[32];
    constexpr int V = Vec4<T>::N;
    typedef typename Vec4<T>::type VT;
    const bool vec = ((Tn % V) == 0) && ((((uintptr_t)x | (uintptr_t)y) & 15) == 0);
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int s = (int)(row % S);
        const T *xr = x + row * (long long)Tn;
        T *yr = y + row * (long long)Tn;
        // pass 1: sum over [htr1, htrn) - the row stays in L1/L2 for pass 2
        double acc = 0.0;
        if (vec) {
            const int c0 = (htr1 / V) * V;
            for (int c = c0 + threadIdx.x * V; c < htrn; c += blockDim.x * V) {
                VT a = *reinterpret_cast<const VT *>(xr + c);
                const T *e = reinterpret_cast<const T *>(&a);
#pragma unroll
                for (int j = 0; j < V; ++j)
                    if (c + j >= htr1 && c + j < htrn) acc += (double)e[j];
            }
        } else {
            for (int c = htr1 + threadIdx.x; c < htrn; c += blockDim.x) acc += (double)xr[c];
        }
        const double total = block_sum(acc, red);
        double avg = total / (double)(htrn - htr1) * taper[s];
        if (trunc_avg) avg = trunc(avg);
        const T avg_t = (T)avg;  // np.atleast_2d(avg_trace).transpose().astype(self.data.dtype)
        if (vec) {
            for (int c = threadIdx.x * V; c < Tn; c += blockDim.x * V) {
                VT a = *reinterpret_cast<const VT *>(xr + c);
                T *e = reinterpret_cast<T *>(&a);
#pragma unroll
                for (int j = 0; j < V; ++j) e[j] = e[j] - avg_t;
                *reinterpret_cast<VT *>(yr + c) = a;
            }
        } else {
            for (int c = threadIdx.x; c < Tn; c += blockDim.x) yr[c] = xr[c] - avg_t;
        }
    }
}

// ------------------------------------------------------------------------------------------- ahfilt
// Column window of trace i (python slice semantics), _RadarDataFiltering.py:67-72.
__device__ __forceinline__ void ahfilt_window(int i, int Tn, int w, int tail_lo, int &lo, int &hi) {
    const int h = w / 2;
    if (i <= h) {
        lo = 0;
        hi = min(h + i, Tn);
    } else if (i >= Tn - h) {
        lo = tail_lo;
        hi = Tn;
    } else {
        lo = i - h + 1;
        hi = i + h;
    }
    if (hi < lo) hi = lo;
}

template <typename T>
__global__ void __launch_bounds__(256) ahfilt_kernel(const T *__restrict__ x, T *__restrict__ y, int S, int Tn,
                                                     long long rows, int w, int tail_lo,
                                                     const double *__restrict__ taper,
                                                     double *__restrict__ gscratch) {
    extern __shared__ double smem_d[];
    __shared__ double warp_tot[8];
    __shared__ double carry_s;
    double *P = gscratch ? gscratch + (size_t)blockIdx.x * (size_t)(Tn + 1) : smem_d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const long long b = row / S;
        const int s = (int)(row % S);
        const T *xb = x + b * (long long)S * Tn;
        // the 7-tap triangular kernel on the odd-extended trace, folded onto real rows
        int rrow[14];
        double rcoef[14];
        int nr = 0;
#pragma unroll
        for (int j = -3; j <= 3; ++j) {
            const double c = (double)(4 - (j < 0 ? -j : j)) / 16.0;
            const int n = s + j;
            if (n < 0) {
                rrow[nr] = 0; rcoef[nr++] = 2.0 * c;
                rrow[nr] = -n; rcoef[nr++] = -c;
            } else if (n >= S) {
                rrow[nr] = S - 1; rcoef[nr++] = 2.0 * c;
                rrow[nr] = 2 * (S - 1) - n; rcoef[nr++] = -c;
            } else {
                rrow[nr] = n; rcoef[nr++] = c;
            }
        }
        if (threadIdx.x == 0) {
            carry_s = 0.0;
            P[0] = 0.0;
        }
        __syncthreads();
        // inclusive prefix sum of f[c] = sum_j coef_j * x[row_j][c], chunk of 256 columns at a time
        for (int c0 = 0; c0 < Tn; c0 += blockDim.x) {
            const int c = c0 + threadIdx.x;
            double f = 0.0;
            if (c < Tn) {
                for (int j = 0; j < nr; ++j) f += rcoef[j] * (double)xb[(long long)rrow[j] * Tn + c];
            }
            // warp inclusive scan
            double v = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double n = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += n;
            }
            if (lane == 31) warp_tot[warp] = v;
            __syncthreads();
            double off = carry_s;
            for (int k = 0; k < warp; ++k) off += warp_tot[k];
            if (c < Tn) P[c + 1] = v + off;
            __syncthreads();
            if (threadIdx.x == blockDim.x - 1) carry_s = v + off;
            __syncthreads();
        }
        const double tp = taper[s];
        const T *xr = xb + (long long)s * Tn;
        T *yr = y + row * (long long)Tn;
        for (int c = threadIdx.x; c < Tn; c += blockDim.x) {
            int lo, hi;
            ahfilt_window(c, Tn, w, tail_lo, lo, hi);
            const double mean = (P[hi] - P[lo]) / (double)(hi - lo);  // 0/0 -> NaN like np.mean of empty
            yr[c] = (T)((double)xr[c] - mean * tp);
        }
        __syncthreads();
    }
}

// Strip kernel: a CTA owns a strip of W output traces and a chunk of R sample rows.  Rows are visited in time
// order; per row the window means of the strip (fp64 prefix sums over the row segment the windows cover: 16
// consecutive columns per thread, then a block scan of the thread totals) go into a ring of the last seven mean
// rows, and as soon as a row's 7-tap neighbourhood is complete it is written out - one read of the data (plus
// w / W halo) and one write.  Requires n = segment length <= 4096 (256 threads x 16).
constexpr int AH_E = 16;                                  // columns per thread in the scan
__device__ __forceinline__ int ah_pad(int k) { return k + (k >> 4); }   // 17-double pitch: conflict-free chunks

template <typename T>
__global__ void __launch_bounds__(256) ahfilt_strip_kernel(const T *__restrict__ x, T *__restrict__ y, int S, int Tn,
                                                           int w, int tail_lo, const double *__restrict__ taper, int W,
                                                           int R) {
    extern __shared__ double smem_d[];
    double *P = smem_d;                 // [ah_pad(4096)]: inclusive prefix within each thread's chunk
    double *toff = P + 4096 + 256;      // [256] exclusive offsets of the chunks
    double *wtot = toff + 256;          // [8]
    T *ring = reinterpret_cast<T *>(wtot + 8);  // [7][W]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = blockIdx.x * W, c1 = min(Tn, c0 + W);
    const T *xb = x + (long long)blockIdx.z * S * Tn;
    T *yb = y + (long long)blockIdx.z * S * Tn;
    const int h = w / 2;
    int L0, L1;
    {
        int lo, hi;
        ahfilt_window(c0, Tn, w, tail_lo, lo, hi);
        L0 = lo;
        L1 = hi;
        ahfilt_window(c1 - 1, Tn, w, tail_lo, lo, hi);
        L0 = min(L0, lo);
        L1 = max(L1, hi);
        if (c1 - 1 >= Tn - h) {
            L0 = min(L0, tail_lo);
            L1 = Tn;
        }
    }
    const int n = L1 - L0;
    const int s_begin = blockIdx.y * R, s_end = min(S, s_begin + R);
    const int r0 = max(0, s_begin - 3), r1 = min(S - 1, s_end + 2);
    int next_out = s_begin;

    // software pipeline: the row segment of iteration r + 1 and the centre row of this iteration's output are in
    // flight while row r is scanned
    constexpr int NV = 4096 / 256;
    T nxt[NV];
    {
        const T *xr = xb + (long long)r0 * Tn + L0;
#pragma unroll
        for (int u = 0; u < NV; ++u) nxt[u] = (tid + 256 * u < n) ? xr[tid + 256 * u] : (T)0;
    }
    const int nxc = (c1 - c0 + 255) / 256;  // <= 8 centre-row values per thread
    // per-column window constants (the same for every row): padded prefix indices, chunk ids, 1 / width
    int wa[8], we[8];
    double winv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int i = c0 + tid + 256 * u;
        int lo = 0, hi = 0;
        if (i < c1) ahfilt_window(i, Tn, w, tail_lo, lo, hi);
        wa[u] = lo - L0 - 1;
        we[u] = hi - L0 - 1;
        winv[u] = 1.0 / (double)(hi - lo);   // inf for an empty window: 0 * inf = NaN like np.mean of an empty slice
    }
    for (int r = r0; r <= r1; ++r) {
        // stage the row segment (coalesced) as doubles
#pragma unroll
        for (int u = 0; u < NV; ++u)
            if (tid + 256 * u < n) P[ah_pad(tid + 256 * u)] = (double)nxt[u];
        if (r < r1) {
            const T *xr = xb + (long long)(r + 1) * Tn + L0;
#pragma unroll
            for (int u = 0; u < NV; ++u) nxt[u] = (tid + 256 * u < n) ? xr[tid + 256 * u] : (T)0;
        }
        T xc[8];
        {
            const int s = next_out;  // the row this iteration will (normally) emit
            const T *xs = xb + (long long)min(s, S - 1) * Tn + c0;
#pragma unroll
            for (int u = 0; u < 8; ++u) xc[u] = (u < nxc && c0 + tid + 256 * u < c1) ? xs[tid + 256 * u] : (T)0;
        }
        __syncthreads();
        // chunk-local inclusive scan: thread t owns columns [16 t, 16 t + 16)
        double tot = 0.0;
        {
            double v[AH_E];
            double *pc = P + 17 * tid;
            const int kb = AH_E * tid;
#pragma unroll
            for (int e = 0; e < AH_E; ++e) v[e] = (kb + e < n) ? pc[e] : 0.0;
#pragma unroll
            for (int e = 0; e < AH_E; ++e) {
                tot += v[e];
                v[e] = tot;
            }
            if (kb < n) {
#pragma unroll
                for (int e = 0; e < AH_E; ++e) pc[e] = v[e];
            }
        }
        // exclusive scan of the 256 chunk totals
        double inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        {
            double wo = 0.0;
            for (int u = 0; u < warp; ++u) wo += wtot[u];
            toff[tid] = wo + inc - tot;
        }
        __syncthreads();
        T *mrow = ring + (r % 7) * W;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = tid + 256 * u;
            if (c0 + c < c1) {
                const int a = wa[u], e = we[u];
                const double pa = (a < 0) ? 0.0 : P[ah_pad(a)] + toff[a >> 4];
                const double pe = (e < 0) ? 0.0 : P[ah_pad(e)] + toff[e >> 4];
                mrow[c] = (T)((pe - pa) * winv[u]);
            }
        }
        __syncthreads();
        bool first_emit = true;
        while (next_out < s_end && (next_out + 3 <= r || r == S - 1)) {
            const int s = next_out++;
            const bool have_xc = first_emit;  // xc holds row s only for the first row emitted in this iteration
            first_emit = false;
            const double tp = taper[s];
            const T *xs = xb + (long long)s * Tn;
            T *ys = yb + (long long)s * Tn;
            if (s >= 3 && s + 3 < S) {
                // interior: the 7-tap triangular kernel [1 2 3 4 3 2 1] / 16 on rows s-3 .. s+3
                const T *q0 = ring + ((s - 3) % 7) * W, *q1 = ring + ((s - 2) % 7) * W, *q2 = ring + ((s - 1) % 7) * W;
                const T *q3 = ring + (s % 7) * W, *q4 = ring + ((s + 1) % 7) * W, *q5 = ring + ((s + 2) % 7) * W;
                const T *q6 = ring + ((s + 3) % 7) * W;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = tid + 256 * u, i = c0 + c;
                    if (i < c1) {
                        // means are stored in T; for float profiles the 7-tap combination also runs in fp32
                        const T f = (T)0.0625 * (q0[c] + q6[c]) + (T)0.125 * (q1[c] + q5[c]) + (T)0.1875 * (q2[c] + q4[c]) +
                                    (T)0.25 * q3[c];
                        const T xv = have_xc ? xc[u] : xs[i];
                        ys[i] = xv - f * (T)tp;
                    }
                }
            } else {
                // edges: the kernel on the odd-extended mean trace, folded onto real rows
                for (int i = c0 + tid; i < c1; i += 256) {
                    const int c = i - c0;
                    double f = 0.0;
#pragma unroll
                    for (int j = -3; j <= 3; ++j) {
                        const double cj = (double)(4 - (j < 0 ? -j : j)) / 16.0;
                        const int q = s + j;
                        if (q < 0) f += cj * (2.0 * (double)ring[c] - (double)ring[((-q) % 7) * W + c]);
                        else if (q >= S) f += cj * (2.0 * (double)ring[((S - 1) % 7) * W + c] -
                                                    (double)ring[((2 * (S - 1) - q) % 7) * W + c]);
                        else f += cj * (double)ring[(q % 7) * W + c];
                    }
                    ys[i] = (T)((double)xs[i] - f * tp);
                }
            }
        }
    }
}

// Fast strip kernel (float radargrams with tnum % 4 == 0 - every BASELINE shape; the default).  Same decomposition as
// ahfilt_strip_kernel - strip of W output traces x chunk of R rows per CTA, float64 prefix sums of the row segment the
// strip's windows cover, ring of the last seven mean rows, one read and one write of the data - rebuilt around the
// instruction count, which is what bounded the first version (131 thread-instructions per sample, 56 % issue
// utilisation, profiles/r01g_ncu_full_ahfilt.txt):
//   * a thread owns 16 CONTIGUOUS segment elements, loaded as four aligned float4 straight into registers (the segment
//     start is rounded down to a multiple of 4), scanned in registers, and written once to shared memory as the finished
//     exclusive prefix (offset folded in) - no staging round trip, no second pass over the prefix buffer;
//   * the window of every output column is row-invariant: its two padded prefix indices and 1 / width are computed once
//     per CTA and kept in registers, so a window mean is two LDS.64, a DSUB and a DMUL - no per-element window logic;
//   * a thread owns PAIRS of adjacent output columns: ring rows, x and y move as float2 and the 7-tap combination runs
//     on the packed fp32x2 pipe.
constexpr int AHF_E = 16;                 // segment elements per thread
constexpr int AHF_NMAX = 256 * AHF_E;     // longest row segment (strip + window halo)
constexpr int AHF_PAIRS = 4;              // column pairs per thread -> strips of up to 2048 output traces
__device__ __forceinline__ int ahf_pad(int k) { return k + (k >> 4); }   // 17-double pitch per thread: conflict-free

__global__ void __launch_bounds__(256, 2) ahfilt_fast_kernel(const float *__restrict__ x, float *__restrict__ y, int S,
                                                             int Tn, int w, int tail_lo,
                                                             const double *__restrict__ taper, int W, int R) {
    extern __shared__ double smem_d[];
    double *Q = smem_d;                                   // exclusive prefix: Q[ahf_pad(k)] = sum of segment elements < k
    double *wtot = Q + (AHF_NMAX + AHF_NMAX / 16 + 8);    // [8]
    float *ring = reinterpret_cast<float *>(wtot + 8);    // [7][W]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = blockIdx.x * W, c1 = min(Tn, c0 + W);
    const float *xb = x + (long long)blockIdx.z * S * Tn;
    float *yb = y + (long long)blockIdx.z * S * Tn;
    const int h = w / 2;
    int L0, L1;
    {
        int lo, hi;
        ahfilt_window(c0, Tn, w, tail_lo, lo, hi);
        L0 = lo;
        L1 = hi;
        ahfilt_window(c1 - 1, Tn, w, tail_lo, lo, hi);
        L0 = min(L0, lo);
        L1 = max(L1, hi);
        if (c1 - 1 >= Tn - h) {
            L0 = min(L0, tail_lo);
            L1 = Tn;
        }
    }
    L0 &= ~3;                                             // aligned float4 loads (tnum % 4 == 0, 16-byte aligned rows)
    const int s_begin = blockIdx.y * R, s_end = min(S, s_begin + R);
    const int r0 = max(0, s_begin - 3), r1 = min(S - 1, s_end + 2);
    int next_out = s_begin;

    // row-invariant window constants of this thread's eight output columns
    int ia[2 * AHF_PAIRS], ie[2 * AHF_PAIRS];
    double inv[2 * AHF_PAIRS];
#pragma unroll
    for (int u = 0; u < 2 * AHF_PAIRS; ++u) {
        const int i = c0 + 2 * tid + 512 * (u >> 1) + (u & 1);
        int lo = L0, hi = L0;
        if (i < c1) ahfilt_window(i, Tn, w, tail_lo, lo, hi);
        ia[u] = ahf_pad(lo - L0);
        ie[u] = ahf_pad(hi - L0);
        inv[u] = 1.0 / (double)(hi - lo);   // inf for an empty window: 0 * inf = NaN like np.mean of an empty slice
    }
    // software pipeline: the row segment of iteration r + 1 and the centre row of this iteration's output are in flight
    // while row r is scanned
    const int kb = AHF_E * tid;             // this thread's first segment element
    float4 nxt[AHF_E / 4];
    auto load_segment = [&](int r) {
        const float *xr = xb + (long long)r * Tn + L0 + kb;
#pragma unroll
        for (int v4 = 0; v4 < AHF_E / 4; ++v4)
            nxt[v4] = (L0 + kb + 4 * v4 < Tn) ? *reinterpret_cast<const float4 *>(xr + 4 * v4) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    load_segment(r0);
    if (tid == 0) Q[0] = 0.0;
    for (int r = r0; r <= r1; ++r) {
        // ---- thread-local inclusive scan of 16 contiguous elements, in float64 registers
        // (four independent chains of four, then the chain offsets: depth 6 instead of 16 dependent float64 adds)
        double v[AHF_E];
#pragma unroll
        for (int v4 = 0; v4 < AHF_E / 4; ++v4) {
            v[4 * v4 + 0] = (double)nxt[v4].x;
            v[4 * v4 + 1] = v[4 * v4 + 0] + (double)nxt[v4].y;
            v[4 * v4 + 2] = v[4 * v4 + 1] + (double)nxt[v4].z;
            v[4 * v4 + 3] = v[4 * v4 + 2] + (double)nxt[v4].w;
        }
        {
            const double o1 = v[3], o2 = o1 + v[7], o3 = o2 + v[11];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                v[4 + e] += o1;
                v[8 + e] += o2;
                v[12 + e] += o3;
            }
        }
        const double run = v[AHF_E - 1];
        if (r < r1) load_segment(r + 1);
        float2 xc[AHF_PAIRS];
        {
            const float *xs = xb + (long long)min(next_out, S - 1) * Tn + c0 + 2 * tid;
#pragma unroll
            for (int u = 0; u < AHF_PAIRS; ++u)
                xc[u] = (c0 + 2 * tid + 512 * u < c1) ? *reinterpret_cast<const float2 *>(xs + 512 * u) : make_float2(0.f, 0.f);
        }
        // ---- exclusive offset of this thread's chunk: warp scan of the chunk totals, then the warp totals
        double inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();                     // also: the previous row's window reads of Q and ring reads are done
        double off = inc - run;
        {
            // exclusive scan of the eight warp totals: every warp does it redundantly in its first eight lanes
            double wv = (lane < 8) ? wtot[lane] : 0.0;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, wv, o);
                if (lane >= o) wv += up;
            }
            const double wprev = __shfl_sync(0xffffffffu, wv, max(warp - 1, 0));
            if (warp > 0) off += wprev;
        }
        {
            double *qc = Q + 17 * tid + 1;   // ahf_pad(16 tid + 1 + e) = 17 tid + 1 + e for e < 15, + 1 more for e = 15
#pragma unroll
            for (int e = 0; e < AHF_E - 1; ++e) qc[e] = off + v[e];
            qc[AHF_E] = off + v[AHF_E - 1];
        }
        __syncthreads();
        // ---- window means of this thread's column pairs into the ring
        float *mrow = ring + (r % 7) * W + 2 * tid;
#pragma unroll
        for (int u = 0; u < AHF_PAIRS; ++u) {
            if (c0 + 2 * tid + 512 * u < c1) {
                const float m0 = (float)((Q[ie[2 * u]] - Q[ia[2 * u]]) * inv[2 * u]);
                const float m1 = (float)((Q[ie[2 * u + 1]] - Q[ia[2 * u + 1]]) * inv[2 * u + 1]);
                *reinterpret_cast<float2 *>(mrow + 512 * u) = make_float2(m0, m1);
            }
        }
        __syncthreads();
        bool first_emit = true;
        while (next_out < s_end && (next_out + 3 <= r || r == S - 1)) {
            const int s = next_out++;
            const bool have_xc = first_emit;  // xc holds row s only for the first row emitted in this iteration
            first_emit = false;
            const double tp = taper[s];
            const float *xs = xb + (long long)s * Tn;
            float *ys = yb + (long long)s * Tn;
            if (s >= 3 && s + 3 < S) {
                // interior: the 7-tap triangular kernel [1 2 3 4 3 2 1] / 16 on rows s-3 .. s+3, two columns at a time
                const float *q0 = ring + ((s + 4) % 7) * W + 2 * tid, *q1 = ring + ((s + 5) % 7) * W + 2 * tid;
                const float *q2 = ring + ((s + 6) % 7) * W + 2 * tid, *q3 = ring + (s % 7) * W + 2 * tid;
                const float *q4 = ring + ((s + 1) % 7) * W + 2 * tid, *q5 = ring + ((s + 2) % 7) * W + 2 * tid;
                const float *q6 = ring + ((s + 3) % 7) * W + 2 * tid;
                const float tpf = (float)tp;
                const float2 ntp = make_float2(-tpf, -tpf);
#pragma unroll
                for (int u = 0; u < AHF_PAIRS; ++u) {
                    const int i = c0 + 2 * tid + 512 * u;
                    if (i < c1) {
                        const int o = 512 * u;
                        const float2 a06 = __fadd2_rn(*reinterpret_cast<const float2 *>(q0 + o), *reinterpret_cast<const float2 *>(q6 + o));
                        const float2 a15 = __fadd2_rn(*reinterpret_cast<const float2 *>(q1 + o), *reinterpret_cast<const float2 *>(q5 + o));
                        const float2 a24 = __fadd2_rn(*reinterpret_cast<const float2 *>(q2 + o), *reinterpret_cast<const float2 *>(q4 + o));
                        float2 f = __fmul2_rn(make_float2(0.0625f, 0.0625f), a06);
                        f = __ffma2_rn(make_float2(0.125f, 0.125f), a15, f);
                        f = __ffma2_rn(make_float2(0.1875f, 0.1875f), a24, f);
                        f = __ffma2_rn(make_float2(0.25f, 0.25f), *reinterpret_cast<const float2 *>(q3 + o), f);
                        const float2 xv = have_xc ? xc[u] : *reinterpret_cast<const float2 *>(xs + i);
                        *reinterpret_cast<float2 *>(ys + i) = __ffma2_rn(f, ntp, xv);       // x - f * taper
                    }
                }
            } else {
                // edges: the kernel on the odd-extended mean trace, folded onto real rows
                for (int i = c0 + tid; i < c1; i += 256) {
                    const int c = i - c0;
                    double f = 0.0;
#pragma unroll
                    for (int j = -3; j <= 3; ++j) {
                        const double cj = (double)(4 - (j < 0 ? -j : j)) / 16.0;
                        const int q = s + j;
                        if (q < 0) f += cj * (2.0 * (double)ring[c] - (double)ring[((-q) % 7) * W + c]);
                        else if (q >= S) f += cj * (2.0 * (double)ring[((S - 1) % 7) * W + c] -
                                                    (double)ring[((2 * (S - 1) - q) % 7) * W + c]);
                        else f += cj * (double)ring[(q % 7) * W + c];
                    }
                    ys[i] = (float)((double)xs[i] - f * tp);
                }
            }
        }
    }
}

// Sliding-window kernel (for windows too wide for the strip kernel's prefix buffer; testing hook 3): one WARP owns a strip of WS output traces and a chunk of R sample rows; no block-level
// synchronisation at all.  Per row the window sum of the strip's first trace is reduced cooperatively, then the sums of
// the following traces come from the differences sum(i) - sum(i-1) = (elements entering) - (elements leaving) - in the
// interior exactly x[i+h-1] - x[i-h] - accumulated by a float64 warp scan, 64 traces per step (two independent scans in
// flight) with the running sum carried in a register.  Window means go into the warp's own ring of the last seven mean
// rows in shared memory; a row is written out as soon as its 7-tap neighbourhood is complete.  Measured on B200
// (profiles/r01g_ahfilt_slide_*.txt): 1.13-1.20 ms for 8 x 2048 x 8192 at w = 1000 against 0.90 ms for the strip kernel -
// the strip-start window sum (w / 32 loads per row and warp) and the 40 double shuffles per 256 traces cost as much as
// the strip kernel's block-wide prefix sums - so the strip kernel stays the default where it applies.
template <typename T>
__device__ __forceinline__ double ah_delta(const T *__restrict__ xr, int i, int Tn, int w, int tail_lo) {
    int lo, hi, lp, hp;
    ahfilt_window(i, Tn, w, tail_lo, lo, hi);
    ahfilt_window(i - 1, Tn, w, tail_lo, lp, hp);
    double d = 0.0;
    for (int k = hp; k < hi; ++k) d += (double)xr[k];
    for (int k = hi; k < hp; ++k) d -= (double)xr[k];
    for (int k = lp; k < lo; ++k) d -= (double)xr[k];
    for (int k = lo; k < lp; ++k) d += (double)xr[k];
    return d;
}

__device__ __forceinline__ double warp_scan_incl(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

template <typename T, int NE>   // strip width WS = 32 * NE traces; lane l owns traces c0 + 32 e + l, e < NE
__global__ void __launch_bounds__(256) ahfilt_slide_kernel(const T *__restrict__ x, T *__restrict__ y, int S, int Tn,
                                                           int w, int tail_lo, const double *__restrict__ taper, int R) {
    constexpr int WS = 32 * NE;
    extern __shared__ double smem_d[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T *ring = reinterpret_cast<T *>(smem_d) + (size_t)warp * 7 * WS;   // [7][WS], this warp's own
    const int c0 = (blockIdx.x * nw + warp) * WS;
    if (c0 >= Tn) return;
    const int c1 = min(Tn, c0 + WS);
    const T *xb = x + (long long)blockIdx.z * S * Tn;
    T *yb = y + (long long)blockIdx.z * S * Tn;
    const int s_begin = blockIdx.y * R, s_end = min(S, s_begin + R);
    const int r0 = max(0, s_begin - 3), r1 = min(S - 1, s_end + 2);
    int next_out = s_begin;
    int lo0, hi0;
    ahfilt_window(c0, Tn, w, tail_lo, lo0, hi0);
    // Row-invariant description of every trace this lane owns: sum(i) - sum(i-1) = x[ent] - x[lev] (either may be absent:
    // -1); the few traces where a window bound moves by more than one element (regime changes) take the general path.
    int ent[NE], lev[NE];
    double wv[NE];
    unsigned rare = 0;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int i = c0 + 32 * e + lane;
        ent[e] = lev[e] = -1;
        wv[e] = 0.0;
        if (i < c1) {
            int lo, hi, lp, hp;
            ahfilt_window(i, Tn, w, tail_lo, lo, hi);
            wv[e] = __drcp_rn((double)(hi - lo));   // inf for an empty window: 0 * inf = NaN like np.mean of an empty slice
            if (i > c0) {
                ahfilt_window(i - 1, Tn, w, tail_lo, lp, hp);
                if (hi == hp + 1) ent[e] = hp;
                else if (hi != hp) rare |= 1u << e;
                if (lo == lp + 1) lev[e] = lp;
                else if (lo != lp) rare |= 1u << e;
            }
        }
    }
    const bool any_rare = __any_sync(0xffffffffu, rare != 0);

    for (int r = r0; r <= r1; ++r) {
        const T *xr = xb + (long long)r * Tn;
        // ---- every load of this row first (independent: they overlap), then the arithmetic
        T vin[NE], vout[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            vin[e] = ent[e] >= 0 ? xr[ent[e]] : (T)0;
            vout[e] = lev[e] >= 0 ? xr[lev[e]] : (T)0;
        }
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
        {
            int k = lo0 + lane;
            for (; k + 96 < hi0; k += 128) {
                const T a0 = xr[k], a1 = xr[k + 32], a2 = xr[k + 64], a3 = xr[k + 96];
                p0 += (double)a0; p1 += (double)a1; p2 += (double)a2; p3 += (double)a3;
            }
            for (; k < hi0; k += 32) p0 += (double)xr[k];
        }
        double d[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) d[e] = (double)vin[e] - (double)vout[e];
        if (any_rare) {
#pragma unroll
            for (int e = 0; e < NE; ++e)
                if (rare & (1u << e)) d[e] = ah_delta(xr, c0 + 32 * e + lane, Tn, w, tail_lo);
        }
        // window sum of the strip's first trace: the scan starts from it
        const double s0 = warp_sum((p0 + p1) + (p2 + p3));
        if (lane == 0) d[0] = s0;
        // ---- NE independent warp scans (interleaved by the unroll), then the strip-wide offsets
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const double u = __shfl_up_sync(0xffffffffu, d[e], o);
                if (lane >= o) d[e] += u;
            }
        }
        T *mrow = ring + (r % 7) * WS;
        double carry = 0.0;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const double tot = __shfl_sync(0xffffffffu, d[e], 31);
            const double sum = carry + d[e];
            carry += tot;
            if (c0 + 32 * e + lane < c1) mrow[32 * e + lane] = (T)(sum * wv[e]);
        }
        __syncwarp();
        while (next_out < s_end && (next_out + 3 <= r || r == S - 1)) {
            const int s = next_out++;
            const double tp = taper[s];
            const T *xs = xb + (long long)s * Tn + c0;
            T *ys = yb + (long long)s * Tn + c0;
            T xv[NE];
#pragma unroll
            for (int e = 0; e < NE; ++e) xv[e] = (c0 + 32 * e + lane < c1) ? xs[32 * e + lane] : (T)0;
            if (s >= 3 && s + 3 < S) {
                const T *q0 = ring + ((s - 3) % 7) * WS, *q1 = ring + ((s - 2) % 7) * WS, *q2 = ring + ((s - 1) % 7) * WS;
                const T *q3 = ring + (s % 7) * WS, *q4 = ring + ((s + 1) % 7) * WS, *q5 = ring + ((s + 2) % 7) * WS;
                const T *q6 = ring + ((s + 3) % 7) * WS;
                const T tpT = (T)tp;
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const int c = 32 * e + lane;
                    const T f = (T)0.0625 * (q0[c] + q6[c]) + (T)0.125 * (q1[c] + q5[c]) + (T)0.1875 * (q2[c] + q4[c]) +
                                (T)0.25 * q3[c];
                    if (c0 + c < c1) ys[c] = xv[e] - f * tpT;
                }
            } else {
                // edges: the kernel on the odd-extended mean trace, folded onto real rows
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const int c = 32 * e + lane;
                    if (c0 + c >= c1) continue;
                    double f = 0.0;
#pragma unroll
                    for (int j = -3; j <= 3; ++j) {
                        const double cj = (double)(4 - (j < 0 ? -j : j)) / 16.0;
                        const int q = s + j;
                        if (q < 0) f += cj * (2.0 * (double)ring[c] - (double)ring[((-q) % 7) * WS + c]);
                        else if (q >= S) f += cj * (2.0 * (double)ring[((S - 1) % 7) * WS + c] -
                                                    (double)ring[((2 * (S - 1) - q) % 7) * WS + c]);
                        else f += cj * (double)ring[(q % 7) * WS + c];
                    }
                    ys[c] = (T)((double)xv[e] - f * tp);
                }
            }
        }
        __syncwarp();
    }
}

template <typename T, int NE>
static int launch_ahfilt_slide(const T *x, T *y, int S, int Tn, int batch, int w, int tail_lo, const double *taper, int R,
                               cudaStream_t st) {
    constexpr int WS = 32 * NE;
    const int nstrips = (Tn + WS - 1) / WS;
    const int nw = nstrips < 8 ? nstrips : 8;
    const size_t smem = (size_t)nw * 7 * WS * sizeof(T);
    static bool attr_done_dev[IMPDAR_MAX_DEVICES];
    bool &attr_done = attr_done_dev[current_device_slot()];
    if (!attr_done) {
        if (8 * 7 * WS * sizeof(T) > 48 * 1024)
            IMPDAR_CUDA(cudaFuncSetAttribute(ahfilt_slide_kernel<T, NE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(8 * 7 * WS * sizeof(T))));
        attr_done = true;
    }
    dim3 grid((nstrips + nw - 1) / nw, (S + R - 1) / R, batch);
    IMPDAR_CHECK_ARG(batch <= 65535 && grid.y <= 65535, "ahfilt: batch too large");
    ktimer_begin("ahfilt_slide_kernel", st);
    ahfilt_slide_kernel<T, NE><<<grid, nw * 32, smem, st>>>(x, y, S, Tn, w, tail_lo, taper, R);
    ktimer_end(st);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

// ------------------------------------------------------------------------------------------ filtfilt
// Software-pipelined recurrence over n strided samples: the next block of FF_U inputs is in flight while the current
// block runs through the (serial, fp64-pipe bound) recurrence, so the HBM latency hides behind 21 DFMA per sample.
#ifndef IMPDAR_FF_U
#define IMPDAR_FF_U 16
#endif
constexpr int FF_U = IMPDAR_FF_U;
// Optional L2 prefetch distance in blocks of FF_U rows (prefetch.global.L2 of the rows FF_PD blocks ahead).  Measured on
// B200 (profiles/r01g_filtfilt_prefetch_ab.txt): 0.529 ms without, 0.567 / 0.568 / 0.604 ms with FF_PD = 2 / 4 / 8 for
// 8 x 2048 x 8192 - the extra requests cost more than the L2 hits save, so it is off.
#ifndef IMPDAR_FF_PD
#define IMPDAR_FF_PD 0
#endif
constexpr int FF_PD = IMPDAR_FF_PD;
template <typename T>
__device__ __forceinline__ void prefetch_l2(const T *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// `src` walks n samples with stride `ss` elements (negative = backwards), each mapped through `pre` (the odd
// extension 2 x_edge - x of the pads, or the identity); results go to `dst` with stride `ds` when STORE.
template <typename T, typename BUF, int NS, bool STORE, bool BZ, typename Pre>
__device__ __forceinline__ void iir_run(int n, double (&z)[NS], const IirCoef &c, const T *__restrict__ src,
                                        long long ss, T *__restrict__ dst, long long ds, Pre pre) {
    BUF cur[FF_U], nxt[FF_U];
    int i = 0;
    if (n >= FF_U) {
#pragma unroll
        for (int u = 0; u < FF_U; ++u) cur[u] = pre(src[u * ss]);
        src += FF_U * ss;
        for (; i + 2 * FF_U <= n; i += FF_U) {
#pragma unroll
            for (int u = 0; u < FF_U; ++u) nxt[u] = pre(src[u * ss]);
            if (FF_PD > 0 && i + (2 + FF_PD) * FF_U <= n) {
#pragma unroll
                for (int u = 0; u < FF_U; ++u) prefetch_l2(src + (long long)(FF_PD * FF_U + u) * ss);
            }
            src += FF_U * ss;
#pragma unroll
            for (int u = 0; u < FF_U; ++u) {
                const double yv = iir_step<T, NS, BZ>((double)cur[u], z, c);
                if (STORE) dst[u * ds] = (T)yv;
            }
            if (STORE) dst += FF_U * ds;
#pragma unroll
            for (int u = 0; u < FF_U; ++u) cur[u] = nxt[u];
        }
#pragma unroll
        for (int u = 0; u < FF_U; ++u) {
            const double yv = iir_step<T, NS, BZ>((double)cur[u], z, c);
            if (STORE) dst[u * ds] = (T)yv;
        }
        if (STORE) dst += FF_U * ds;
        i += FF_U;
    }
    for (; i < n; ++i) {
        const double yv = iir_step<T, NS, BZ>((double)pre(*src), z, c);
        src += ss;
        if (STORE) {
            *dst = (T)yv;
            dst += ds;
        }
    }
}

template <typename T, int NS, bool BZ>
__global__ void __launch_bounds__(256) filtfilt_kernel(const T *__restrict__ x, T *__restrict__ y,
                                                      T *__restrict__ work, int S, int Tn, long long ntraces,
                                                      int padlen, const __grid_constant__ IirCoef c) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= ntraces) return;
    const long long b = gid / Tn;
    const int t = (int)(gid % Tn);
    const long long st = Tn;
    const T *xb = x + b * (long long)S * Tn + t;
    T *yb = y + b * (long long)S * Tn + t;
    const int L = S + 2 * padlen;
    T *wk = work + b * (long long)L * Tn + t;

    const double x0 = (double)xb[0];
    const double xl = (double)xb[(long long)(S - 1) * st];
    double z[NS];
    // ---- forward over the odd-extended trace (scipy.signal.filtfilt, padtype='odd')
    {
        const double e0 = (padlen > 0) ? 2.0 * x0 - (double)xb[(long long)padlen * st] : x0;
#pragma unroll
        for (int i = 0; i < NS; ++i) z[i] = c.zi[i] * e0;
    }
    iir_run<T, double, NS, true, BZ>(padlen, z, c, xb + (long long)padlen * st, -st, wk, st,
                                 [&](T v) { return 2.0 * x0 - (double)v; });
    iir_run<T, T, NS, true, BZ>(S, z, c, xb, st, wk + (long long)padlen * st, st, [](T v) { return v; });
    iir_run<T, double, NS, true, BZ>(padlen, z, c, xb + (long long)(S - 2) * st, -st, wk + (long long)(padlen + S) * st, st,
                                 [&](T v) { return 2.0 * xl - (double)v; });
    // ---- backward over the forward output
    {
        const double e0 = (double)wk[(long long)(L - 1) * st];
#pragma unroll
        for (int i = 0; i < NS; ++i) z[i] = c.zi[i] * e0;
    }
    iir_run<T, T, NS, false, BZ>(padlen, z, c, wk + (long long)(L - 1) * st, -st, (T *)nullptr, 0, [](T v) { return v; });
    iir_run<T, T, NS, true, BZ>(S, z, c, wk + (long long)(padlen + S - 1) * st, -st, yb + (long long)(S - 1) * st, -st,
                            [](T v) { return v; });
}

template <typename T, int NS>
static int launch_filtfilt(const T *x, T *y, T *work, int S, int Tn, int batch, int padlen, const IirCoef &c,
                           bool odd_b_zero, cudaStream_t st) {
    const long long ntraces = (long long)batch * Tn;
    // one trace per thread: small problems use one warp per CTA so that every SM gets work
    int block = (ntraces < (long long)num_sms() * 64 * 2) ? 32 : 64;
    if (const char *e = getenv("IMPDAR_FF_BLOCK")) block = atoi(e);   // development A/B switch
    const long long grid = (ntraces + block - 1) / block;
    ktimer_begin("filtfilt_kernel", st);
    if (odd_b_zero) filtfilt_kernel<T, NS, true><<<(unsigned)grid, block, 0, st>>>(x, y, work, S, Tn, ntraces, padlen, c);
    else filtfilt_kernel<T, NS, false><<<(unsigned)grid, block, 0, st>>>(x, y, work, S, Tn, ntraces, padlen, c);
    ktimer_end(st);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

template <typename T>
static int filtfilt_impl(const T *x, T *y, int S, int Tn, int batch, const double *b, const double *a,
                         int ncoef, const double *zi, int padlen, void *ws, size_t ws_bytes, void *stream) {
    IMPDAR_CHECK_ARG(x && y && b && a && zi, "filtfilt: null pointer");
    IMPDAR_CHECK_ARG(S > 0 && Tn > 0 && batch > 0, "filtfilt: bad shape");
    IMPDAR_CHECK_ARG(ncoef >= 2 && ncoef <= 33, "filtfilt: ncoef must be in [2, 33], got %d", ncoef);
    IMPDAR_CHECK_ARG(padlen >= 0 && S > padlen,
                     "The length of the input vector x must be greater than padlen, which is %d.", padlen);
    IMPDAR_CHECK_ARG(a[0] == 1.0, "filtfilt: coefficients must be normalised (a[0] == 1)");
    const size_t need = impdar_filtfilt_workspace_bytes(S, Tn, batch, padlen, (int)sizeof(T));
    IMPDAR_CHECK_ARG(ws && ws_bytes >= need, "filtfilt: workspace too small (%zu < %zu)", ws_bytes, need);
    IirCoef c;
    memset(&c, 0, sizeof(c));
    for (int i = 0; i < ncoef; ++i) {
        c.b[i] = b[i];
        c.a[i] = a[i];
    }
    for (int i = 0; i < ncoef - 1; ++i) c.zi[i] = zi[i];
    cudaStream_t st = (cudaStream_t)stream;
    const int ns = ncoef - 1;
    T *work = (T *)ws;
    bool odd_b_zero = ncoef >= 3;        // band-pass numerators: skip the DFMAs that multiply an exact zero
    for (int i = 1; i < ncoef; i += 2) odd_b_zero = odd_b_zero && (b[i] == 0.0);
#define FF_CASE(N) \
    if (ns <= N) return launch_filtfilt<T, N>(x, y, work, S, Tn, batch, padlen, c, odd_b_zero, st);
    FF_CASE(2) FF_CASE(4) FF_CASE(6) FF_CASE(8) FF_CASE(10) FF_CASE(12) FF_CASE(16) FF_CASE(24) FF_CASE(32)
#undef FF_CASE
    set_error("filtfilt: unsupported order");
    return IMPDAR_B200_EINVAL;
}

// ----------------------------------------------------------------------------------------------- FIR
template <typename T>
__global__ void __launch_bounds__(256) fir_kernel(const T *__restrict__ x, T *__restrict__ y, int S, int Tn,
                                                  long long rows, const double *__restrict__ taps, int ntaps) {
    extern __shared__ double tp[];
    for (int i = threadIdx.x; i < ntaps; i += blockDim.x) tp[i] = taps[i];
    __syncthreads();
    const long long total = rows * (long long)Tn;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long row = g / Tn;
        const int s = (int)(row % S);
        const int kmax = min(ntaps - 1, s);
        double acc = 0.0;
        for (int k = 0; k <= kmax; ++k) acc = fma(tp[k], (double)x[g - (long long)k * Tn], acc);
        y[g] = (T)acc;
    }
}

template <typename T>
static int fir_impl(const T *x, T *y, int S, int Tn, int batch, const double *taps, int ntaps, void *stream) {
    IMPDAR_CHECK_ARG(x && y && taps, "fir: null pointer");
    IMPDAR_CHECK_ARG(S > 0 && Tn > 0 && batch > 0, "fir: bad shape");
    IMPDAR_CHECK_ARG(ntaps >= 1 && ntaps <= 1024, "fir: ntaps must be in [1, 1024]");
    IMPDAR_CHECK_ARG(x != y, "fir: in-place not supported");
    cudaStream_t st = (cudaStream_t)stream;
    double *dtaps = nullptr;
    IMPDAR_CUDA(cudaMallocAsync((void **)&dtaps, ntaps * sizeof(double), st));
    IMPDAR_CUDA(cudaMemcpyAsync(dtaps, taps, ntaps * sizeof(double), cudaMemcpyHostToDevice, st));
    const long long total = (long long)batch * S * Tn;
    long long grid = (total + 255) / 256;
    if (grid > (long long)num_sms() * 16) grid = (long long)num_sms() * 16;
    fir_kernel<T><<<(unsigned)grid, 256, ntaps * sizeof(double), st>>>(x, y, S, Tn, (long long)batch * S, dtaps,
                                                                      ntaps);
    IMPDAR_LAUNCH_CHECK();
    IMPDAR_CUDA(cudaFreeAsync(dtaps, st));
    return IMPDAR_B200_OK;
}

template <typename T>
static int hfilt_impl(const T *x, T *y, int S, int Tn, int batch, int htr1, int htrn, const double *taper,
                      int trunc_avg, void *stream) {
    IMPDAR_CHECK_ARG(x && y && taper, "hfilt: null pointer");
    IMPDAR_CHECK_ARG(S > 0 && Tn > 0 && batch > 0, "hfilt: bad shape");
    IMPDAR_CHECK_ARG(0 <= htr1 && htr1 < htrn && htrn <= Tn, "hfilt: bad trace bounds [%d, %d)", htr1, htrn);
    const long long rows = (long long)batch * S;
    long long grid = rows;
    const long long cap = (long long)num_sms() * 32;
    if (grid > cap) grid = cap;
    ktimer_begin("hfilt_kernel", (cudaStream_t)stream);
    hfilt_kernel<T><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, y, S, Tn, rows, htr1, htrn, taper,
                                                                     trunc_avg);
    ktimer_end((cudaStream_t)stream);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

static const size_t AHFILT_SMEM_LIMIT = 200 * 1024;
static int g_ahfilt_force_rowwise = 0;  // testing hook: 0 = auto (fast strip kernel, strip kernel, else warp-sliding), 1 = one-row-per-CTA kernel, 2 = first strip kernel, 3 = warp-sliding kernel

template <typename T>
static int ahfilt_impl(const T *x, T *y, int S, int Tn, int batch, int w, const double *taper, void *ws,
                       size_t ws_bytes, void *stream) {
    IMPDAR_CHECK_ARG(x && y && taper, "ahfilt: null pointer");
    IMPDAR_CHECK_ARG(S > 0 && Tn > 0 && batch > 0, "ahfilt: bad shape");
    IMPDAR_CHECK_ARG(S > 12, "The length of the input vector x must be greater than padlen, which is 12.");
    IMPDAR_CHECK_ARG(w >= 0, "ahfilt: window_size must be >= 0");
    IMPDAR_CHECK_ARG((const void *)x != (const void *)y, "ahfilt: in-place not supported");
    // python slice start of data[:, tnum - w : tnum]
    int tail_lo = Tn - w;
    if (tail_lo < 0) tail_lo = (tail_lo + Tn < 0) ? 0 : tail_lo + Tn;
    // the warp-sliding kernel: any shape and window; takes the shapes the strip kernel does not cover (hook 3 forces it)
    auto run_slide = [&]() -> int {
        int WS = 256, R = 64;
        while (WS > 64 && WS / 2 >= Tn) WS /= 2;
        auto warps = [&](int ws, int r) { return (long long)batch * ((S + r - 1) / r) * ((Tn + ws - 1) / ws); };
        const long long want = (long long)num_sms() * 16;
        if (warps(WS, R) < want) R = 32;
        if (warps(WS, R) < want && WS > 128) WS /= 2;
        if (warps(WS, R) < want) R = 16;
        if (const char *e = getenv("IMPDAR_AH_WS")) WS = atoi(e);   // development A/B switches
        if (const char *e = getenv("IMPDAR_AH_R")) R = atoi(e);
        cudaStream_t st = (cudaStream_t)stream;
        if (WS >= 256) return launch_ahfilt_slide<T, 8>(x, y, S, Tn, batch, w, tail_lo, taper, R, st);
        if (WS >= 128) return launch_ahfilt_slide<T, 4>(x, y, S, Tn, batch, w, tail_lo, taper, R, st);
        return launch_ahfilt_slide<T, 2>(x, y, S, Tn, batch, w, tail_lo, taper, R, st);
    };
    if (g_ahfilt_force_rowwise == 3) return run_slide();
    // default: the prefix-sum strip kernel, whenever the prefix buffer of one strip (W + ~w columns) and the 7-row ring fit
    // in shared memory
    if (g_ahfilt_force_rowwise != 1) {
        int W = (sizeof(T) == 4) ? 2048 : 1024;
        if (W > Tn) W = ((Tn + 31) / 32) * 32;
        const int h = w / 2;
        int nmax = 0;
        for (int c0 = 0; c0 < Tn; c0 += W) {
            const int c1 = (c0 + W < Tn) ? c0 + W : Tn;
            int lo0, hi0, lo1, hi1;
            auto win = [&](int i, int &lo, int &hi) {
                if (i <= h) { lo = 0; hi = (h + i < Tn) ? h + i : Tn; }
                else if (i >= Tn - h) { lo = tail_lo; hi = Tn; }
                else { lo = i - h + 1; hi = i + h; }
                if (hi < lo) hi = lo;
            };
            win(c0, lo0, hi0);
            win(c1 - 1, lo1, hi1);
            int L0 = lo0 < lo1 ? lo0 : lo1, L1 = hi0 > hi1 ? hi0 : hi1;
            if (c1 - 1 >= Tn - h) { if (tail_lo < L0) L0 = tail_lo; L1 = Tn; }
            if (L1 - L0 > nmax) nmax = L1 - L0;
        }
        const size_t smem_strip = (size_t)(4096 + 256 + 256 + 8) * sizeof(double) + (size_t)7 * W * sizeof(T);
        if (sizeof(T) == 4 && g_ahfilt_force_rowwise == 0 && (Tn & 3) == 0 && (W & 1) == 0 && nmax + 3 <= AHF_NMAX &&
            ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0) {
            int dev = 0;
            IMPDAR_CUDA(cudaGetDevice(&dev));
            static bool fast_attr[64];
            const size_t smem_fast = (size_t)(AHF_NMAX + AHF_NMAX / 16 + 8 + 8) * sizeof(double) + (size_t)7 * W * sizeof(float);
            if (dev >= 0 && dev < 64 && !fast_attr[dev]) {
                IMPDAR_CUDA(cudaFuncSetAttribute(ahfilt_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                fast_attr[dev] = true;
            }
            int R = 64;
            auto ctas = [&](int r) { return (long long)batch * ((S + r - 1) / r) * ((Tn + W - 1) / W); };
            if (ctas(R) < (long long)num_sms() * 4) R = 32;      // single profiles: more row chunks, 3 + 3 halo rows each
            if (const char *e = getenv("IMPDAR_AH_R")) R = atoi(e);   // development A/B switch
            dim3 grid((Tn + W - 1) / W, (S + R - 1) / R, batch);
            IMPDAR_CHECK_ARG(batch <= 65535 && grid.y <= 65535, "ahfilt: batch too large");
            ktimer_begin("ahfilt_fast_kernel", (cudaStream_t)stream);
            ahfilt_fast_kernel<<<grid, 256, smem_fast, (cudaStream_t)stream>>>((const float *)x, (float *)y, S, Tn, w, tail_lo,
                                                                                 taper, W, R);
            ktimer_end((cudaStream_t)stream);
            IMPDAR_LAUNCH_CHECK();
            return IMPDAR_B200_OK;
        }
        if (nmax <= 4096) {
            static bool attr_done_dev[IMPDAR_MAX_DEVICES];
            bool &attr_done = attr_done_dev[current_device_slot()];
            if (!attr_done) {
                IMPDAR_CUDA(cudaFuncSetAttribute(ahfilt_strip_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                attr_done = true;
            }
            const int R = 64;
            dim3 grid((Tn + W - 1) / W, (S + R - 1) / R, batch);
            IMPDAR_CHECK_ARG(batch <= 65535 && grid.y <= 65535, "ahfilt: batch too large");
            ktimer_begin("ahfilt_strip_kernel", (cudaStream_t)stream);
            ahfilt_strip_kernel<T><<<grid, 256, smem_strip, (cudaStream_t)stream>>>(x, y, S, Tn, w, tail_lo, taper, W, R);
            ktimer_end((cudaStream_t)stream);
            IMPDAR_LAUNCH_CHECK();
            return IMPDAR_B200_OK;
        }
        if (g_ahfilt_force_rowwise == 0) return run_slide();   // windows wider than the strip buffer
    }
    const long long rows = (long long)batch * S;
    const size_t need_smem = (size_t)(Tn + 1) * sizeof(double);
    long long grid = rows;
    double *scratch = nullptr;
    size_t smem = need_smem;
    if (need_smem > AHFILT_SMEM_LIMIT) {
        const long long cap = (long long)num_sms() * 4;
        if (grid > cap) grid = cap;
        const size_t need = impdar_ahfilt_workspace_bytes(S, Tn, batch);
        IMPDAR_CHECK_ARG(ws && ws_bytes >= need, "ahfilt: workspace too small (%zu < %zu)", ws_bytes, need);
        scratch = (double *)ws;
        smem = 0;
    } else {
        const long long cap = (long long)num_sms() * 16;
        if (grid > cap) grid = cap;
        if (smem > 48 * 1024)
            IMPDAR_CUDA(cudaFuncSetAttribute(ahfilt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
    }
    ktimer_begin("ahfilt_kernel", (cudaStream_t)stream);
    ahfilt_kernel<T><<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(x, y, S, Tn, rows, w, tail_lo, taper,
                                                                         scratch);
    ktimer_end((cudaStream_t)stream);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

}  // namespace impdar

using namespace impdar;

extern "C" {

int impdar_taper_f32(const float *x, float *y, int S, int T, int batch, double htaper, double vtaper,
                     int trunc_int, void *stream) {
    IMPDAR_CHECK_ARG(x && y, "taper: null pointer");
    IMPDAR_CHECK_ARG(S > 0 && T > 0 && batch > 0, "taper: bad shape");
    const long long rows = (long long)batch * S;
    long long grid = rows;
    const long long cap = (long long)num_sms() * 32;
    if (grid > cap) grid = cap;
    taper_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, y, S, T, rows, htaper, vtaper, trunc_int);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

int impdar_taper_f64(const double *x, double *y, int S, int T, int batch, double htaper, double vtaper, void *stream) {
    IMPDAR_CHECK_ARG(x && y, "taper: null pointer");
    IMPDAR_CHECK_ARG(S > 0 && T > 0 && batch > 0, "taper: bad shape");
    const long long rows = (long long)batch * S;
    long long grid = rows;
    const long long cap = (long long)num_sms() * 32;
    if (grid > cap) grid = cap;
    taper_f64_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, y, S, T, rows, htaper, vtaper);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

int impdar_hfilt_f32(const float *x, float *y, int S, int T, int batch, int htr1, int htrn,
                     const double *taper, int trunc_avg, void *stream) {
    return hfilt_impl<float>(x, y, S, T, batch, htr1, htrn, taper, trunc_avg, stream);
}
int impdar_hfilt_f64(const double *x, double *y, int S, int T, int batch, int htr1, int htrn,
                     const double *taper, int trunc_avg, void *stream) {
    return hfilt_impl<double>(x, y, S, T, batch, htr1, htrn, taper, trunc_avg, stream);
}

size_t impdar_ahfilt_workspace_bytes(int S, int T, int batch) {
    (void)S;
    (void)batch;
    const size_t need_smem = (size_t)(T + 1) * sizeof(double);
    if (need_smem <= AHFILT_SMEM_LIMIT) return 0;
    return (size_t)num_sms() * 4 * (size_t)(T + 1) * sizeof(double);
}
int impdar_ahfilt_force_rowwise(int on) {
    g_ahfilt_force_rowwise = (on >= 0 && on <= 3) ? on : 1;
    return IMPDAR_B200_OK;
}
int impdar_ahfilt_f32(const float *x, float *y, int S, int T, int batch, int w, const double *taper, void *ws,
                      size_t ws_bytes, void *stream) {
    return ahfilt_impl<float>(x, y, S, T, batch, w, taper, ws, ws_bytes, stream);
}
int impdar_ahfilt_f64(const double *x, double *y, int S, int T, int batch, int w, const double *taper, void *ws,
                      size_t ws_bytes, void *stream) {
    return ahfilt_impl<double>(x, y, S, T, batch, w, taper, ws, ws_bytes, stream);
}

size_t impdar_filtfilt_workspace_bytes(int S, int T, int batch, int padlen, int elem_bytes) {
    return (size_t)batch * (size_t)(S + 2 * padlen) * (size_t)T * (size_t)elem_bytes;
}
int impdar_filtfilt_f32(const float *x, float *y, int S, int T, int batch, const double *b, const double *a,
                        int ncoef, const double *zi, int padlen, void *ws, size_t ws_bytes, void *stream) {
    return filtfilt_impl<float>(x, y, S, T, batch, b, a, ncoef, zi, padlen, ws, ws_bytes, stream);
}
int impdar_filtfilt_f64(const double *x, double *y, int S, int T, int batch, const double *b, const double *a,
                        int ncoef, const double *zi, int padlen, void *ws, size_t ws_bytes, void *stream) {
    return filtfilt_impl<double>(x, y, S, T, batch, b, a, ncoef, zi, padlen, ws, ws_bytes, stream);
}
int impdar_fir_f32(const float *x, float *y, int S, int T, int batch, const double *taps, int ntaps,
                   void *stream) {
    return fir_impl<float>(x, y, S, T, batch, taps, ntaps, stream);
}
int impdar_fir_f64(const double *x, double *y, int S, int T, int batch, const double *taps, int ntaps,
                   void *stream) {
    return fir_impl<double>(x, y, S, T, batch, taps, ntaps, stream);
}

}  // extern "C"
