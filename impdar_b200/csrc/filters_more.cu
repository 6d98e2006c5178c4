// Sibling filters of the hot path, on the same building blocks as filters.cu:
//   winavg_hfilt                       RadarData/_RadarDataFiltering.py:353-440   moving-window mean-trace removal
//   highpass / lowpass / horizontal_band_pass   :138-350                          filtfilt along the TRACE axis
//   agc, rangegain                     RadarData/_RadarDataProcessing.py:456-496  per-sample-row gains
// Traces are the contiguous axis.  Row-wise work maps a CTA to a sample row (coalesced 128-bit accesses); the
// trace-axis recurrence maps a lane to a row and stages 32 x 32 tiles through shared memory so that every global
// access is still a full 128-byte line.
#include "common.cuh"
#include "iir.cuh"

namespace impdar {

// ------------------------------------------------------------------------------------------ winavg_hfilt
// out[s, i] = x[s, i] - mean(x[s, max(0, i-h) : min(T, i+h)]) * taper[s],  h = (avg_win - 1) // 2   (:421-434).
// One CTA per sample row: fp64 inclusive prefix sums of the row (shared memory, or a global scratch row for very long
// rows), then one subtraction per trace.  An empty window (h == 0) gives 0/0 = NaN like np.mean of an empty slice.
template <typename T>
__global__ void __launch_bounds__(256) winavg_kernel(const T *__restrict__ x, T *__restrict__ y, int S, int Tn, long long rows,
                                                     int h, const double *__restrict__ taper, double *__restrict__ gscratch) {
    extern __shared__ double smem_d[];
    __shared__ double warp_tot[8];
    __shared__ double carry_s;
    double *P = gscratch ? gscratch + (size_t)blockIdx.x * (size_t)(Tn + 1) : smem_d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int s = (int)(row % S);
        const T *xr = x + row * (long long)Tn;
        T *yr = y + row * (long long)Tn;
        if (threadIdx.x == 0) {
            carry_s = 0.0;
            P[0] = 0.0;
        }
        __syncthreads();
        for (int c0 = 0; c0 < Tn; c0 += blockDim.x) {
            const int c = c0 + threadIdx.x;
            double v = (c < Tn) ? (double)xr[c] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double n = __shfl_up_sync(0xffffffffu, v, o);
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
        for (int c = threadIdx.x; c < Tn; c += blockDim.x) {
            const int lo = max(c - h, 0), hi = min(c + h, Tn);
            const double mean = (P[hi] - P[lo]) / (double)(hi - lo);
            yr[c] = (T)((double)xr[c] - mean * tp);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------- filtfilt along the trace axis
// scipy.signal.filtfilt(b, a, x, axis=-1), padtype 'odd' (:203, :273, :338).  A warp owns 32 consecutive sample rows;
// the odd-extended row (length L = T + 2 padlen) is walked in chunks of 32 samples: the chunk of all 32 rows is loaded
// cooperatively (lane = sample, one 128-byte line per row), transposed through shared memory (lane = row), run through
// the fp64 recurrence, transposed back and stored.  The forward pass is parked in `work` (nrows x L), the backward
// pass reads it in reverse and writes the kept samples.  The next chunk's loads are in flight during the recurrence.
constexpr int FR_PITCH = 33;

template <typename T>
__device__ __forceinline__ double fr_ext(const T *__restrict__ xr, int k, int Tn, int padlen, double x0, double xl) {
    if (k < padlen) return 2.0 * x0 - (double)xr[padlen - k];
    if (k < padlen + Tn) return (double)xr[k - padlen];
    return 2.0 * xl - (double)xr[Tn - 2 - (k - padlen - Tn)];
}

template <typename T, int NS>
__global__ void __launch_bounds__(32) filtfilt_rows_kernel(const T *__restrict__ x, T *__restrict__ y, T *__restrict__ work,
                                                           int Tn, long long nrows, int padlen,
                                                           const __grid_constant__ IirCoef c) {
    __shared__ double tile[32 * FR_PITCH];
    __shared__ double edge0[32], edge1[32];
    const int lane = threadIdx.x;
    const long long r0 = (long long)blockIdx.x * 32;
    const int nr = (int)min((long long)32, nrows - r0);
    const int L = Tn + 2 * padlen;
    if (lane < nr) {
        const T *xr = x + (r0 + lane) * (long long)Tn;
        edge0[lane] = (double)xr[0];
        edge1[lane] = (double)xr[Tn - 1];
    }
    __syncwarp();
    double z[NS];
    double nxt[32];
    // ---------------- forward over the extended rows
    auto load_fwd = [&](int k0) {
        const int k = k0 + lane;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            nxt[i] = (i < nr && k < L) ? fr_ext(x + (r0 + i) * (long long)Tn, k, Tn, padlen, edge0[i], edge1[i]) : 0.0;
    };
    load_fwd(0);
    for (int k0 = 0; k0 < L; k0 += 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) tile[i * FR_PITCH + lane] = nxt[i];
        __syncwarp();
        if (k0 + 32 < L) load_fwd(k0 + 32);
        if (k0 == 0) {
            const double e0 = tile[lane * FR_PITCH];
#pragma unroll
            for (int i = 0; i < NS; ++i) z[i] = c.zi[i] * e0;
        }
        const int nk = min(32, L - k0);
        for (int j = 0; j < nk; ++j) tile[lane * FR_PITCH + j] = iir_step<T, NS>(tile[lane * FR_PITCH + j], z, c);
        __syncwarp();
        if (k0 + lane < L) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < nr) work[(r0 + i) * (long long)L + k0 + lane] = (T)tile[i * FR_PITCH + lane];
        }
        __syncwarp();
    }
    // ---------------- backward over the forward output; chunks aligned to the end of the row
    auto load_bwd = [&](int k0) {
        const int k = k0 + lane;
#pragma unroll
        for (int i = 0; i < 32; ++i) nxt[i] = (i < nr && k >= 0) ? (double)work[(r0 + i) * (long long)L + k] : 0.0;
    };
    load_bwd(L - 32);
    for (int k0 = L - 32; k0 + 32 > padlen; k0 -= 32) {  // the left pad's output is discarded: stop there
#pragma unroll
        for (int i = 0; i < 32; ++i) tile[i * FR_PITCH + lane] = nxt[i];
        __syncwarp();
        if (k0 > padlen) load_bwd(k0 - 32);
        if (k0 == L - 32) {
            const double e0 = tile[lane * FR_PITCH + 31];
#pragma unroll
            for (int i = 0; i < NS; ++i) z[i] = c.zi[i] * e0;
        }
        const int jmin = max(0, -k0);
        for (int j = 31; j >= jmin; --j) tile[lane * FR_PITCH + j] = iir_step<T, NS>(tile[lane * FR_PITCH + j], z, c);
        __syncwarp();
        const int col = k0 + lane - padlen;
        if (col >= 0 && col < Tn) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < nr) y[(r0 + i) * (long long)Tn + col] = (T)tile[i * FR_PITCH + lane];
        }
        __syncwarp();
    }
}

template <typename T, int NS>
static int launch_filtfilt_rows(const T *x, T *y, T *work, int Tn, long long nrows, int padlen, const IirCoef &c,
                                cudaStream_t st) {
    const long long grid = (nrows + 31) / 32;
    ktimer_begin("filtfilt_rows_kernel", st);
    filtfilt_rows_kernel<T, NS><<<(unsigned)grid, 32, 0, st>>>(x, y, work, Tn, nrows, padlen, c);
    ktimer_end(st);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

template <typename T>
static int filtfilt_rows_impl(const T *x, T *y, int S, int Tn, int batch, const double *b, const double *a, int ncoef,
                              const double *zi, int padlen, void *ws, size_t ws_bytes, void *stream) {
    IMPDAR_CHECK_ARG(x && y && b && a && zi, "filtfilt_rows: null pointer");
    IMPDAR_CHECK_ARG(S > 0 && Tn > 0 && batch > 0, "filtfilt_rows: bad shape");
    IMPDAR_CHECK_ARG(ncoef >= 2 && ncoef <= 33, "filtfilt_rows: ncoef must be in [2, 33], got %d", ncoef);
    IMPDAR_CHECK_ARG(padlen >= 0 && Tn > padlen,
                     "The length of the input vector x must be greater than padlen, which is %d.", padlen);
    IMPDAR_CHECK_ARG(a[0] == 1.0, "filtfilt_rows: coefficients must be normalised (a[0] == 1)");
    const size_t need = impdar_filtfilt_rows_workspace_bytes(S, Tn, batch, padlen, (int)sizeof(T));
    IMPDAR_CHECK_ARG(ws && ws_bytes >= need, "filtfilt_rows: workspace too small (%zu < %zu)", ws_bytes, need);
    IirCoef c;
    memset(&c, 0, sizeof(c));
    for (int i = 0; i < ncoef; ++i) {
        c.b[i] = b[i];
        c.a[i] = a[i];
    }
    for (int i = 0; i < ncoef - 1; ++i) c.zi[i] = zi[i];
    cudaStream_t st = (cudaStream_t)stream;
    const int ns = ncoef - 1;
    const long long nrows = (long long)batch * S;
#define FR_CASE(N) \
    if (ns <= N) return launch_filtfilt_rows<T, N>(x, y, (T *)ws, Tn, nrows, padlen, c, st);
    FR_CASE(2) FR_CASE(4) FR_CASE(6) FR_CASE(8) FR_CASE(10) FR_CASE(12) FR_CASE(16) FR_CASE(24) FR_CASE(32)
#undef FR_CASE
    set_error("filtfilt_rows: unsupported order");
    return IMPDAR_B200_EINVAL;
}

// -------------------------------------------------------------------------------------------- row gains
// max_t |x[s, t]| per sample row (agc, _RadarDataProcessing.py:483-485); NaN propagates like np.max.
template <typename T>
__global__ void __launch_bounds__(256) rowabsmax_kernel(const T *__restrict__ x, double *__restrict__ out, int Tn, long long rows) {
    __shared__ double red[8];
    __shared__ int red_nan[8];
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const T *xr = x + row * (long long)Tn;
        double m = 0.0;
        int nan = 0;
        for (int c = threadIdx.x; c < Tn; c += blockDim.x) {
            const double v = fabs((double)xr[c]);
            nan |= (v != v);
            m = fmax(m, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            nan |= __shfl_xor_sync(0xffffffffu, nan, o);
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) {
            red[threadIdx.x >> 5] = m;
            red_nan[threadIdx.x >> 5] = nan;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < 8; ++i) {
                m = fmax(m, red[i]);
                nan |= red_nan[i];
            }
            out[row] = nan ? NAN : m;
        }
    }
}

// y[s, t] = x[s, t] * gain[s] for s > trig[t] (trig == null: every row).  in_double: the product is formed in
// float64 and cast back (numpy's in-place `data *= float64 gain`, rangegain :466-471); otherwise the gain is cast
// to the data type first (agc's `.astype(self.data.dtype)`, :487).
template <typename T>
__global__ void __launch_bounds__(256) rowgain_kernel(const T *__restrict__ x, T *__restrict__ y, int S, int Tn, long long rows,
                                                      const double *__restrict__ gain, const int *__restrict__ trig, int in_double) {
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int s = (int)(row % S);
        const double g = gain[s];
        const T gt = (T)g;
        const T *xr = x + row * (long long)Tn;
        T *yr = y + row * (long long)Tn;
        for (int c = threadIdx.x; c < Tn; c += blockDim.x) {
            const T v = xr[c];
            T r;
            if (trig && s <= trig[c]) r = v;
            else r = in_double ? (T)((double)v * g) : v * gt;
            yr[c] = r;
        }
    }
}

static const size_t WINAVG_SMEM_LIMIT = 200 * 1024;

template <typename T>
static int winavg_impl(const T *x, T *y, int S, int Tn, int batch, int half, const double *taper, void *ws, size_t ws_bytes,
                       void *stream) {
    IMPDAR_CHECK_ARG(x && y && taper, "winavg: null pointer");
    IMPDAR_CHECK_ARG(S > 0 && Tn > 0 && batch > 0, "winavg: bad shape");
    IMPDAR_CHECK_ARG(half >= 0, "winavg: window must be >= 1");
    IMPDAR_CHECK_ARG((const void *)x != (const void *)y, "winavg: in-place not supported");
    const long long rows = (long long)batch * S;
    const size_t need_smem = (size_t)(Tn + 1) * sizeof(double);
    long long grid = rows;
    double *scratch = nullptr;
    size_t smem = need_smem;
    if (need_smem > WINAVG_SMEM_LIMIT) {
        const long long cap = (long long)num_sms() * 4;
        if (grid > cap) grid = cap;
        const size_t need = impdar_winavg_workspace_bytes(S, Tn, batch);
        IMPDAR_CHECK_ARG(ws && ws_bytes >= need, "winavg: workspace too small (%zu < %zu)", ws_bytes, need);
        scratch = (double *)ws;
        smem = 0;
    } else {
        const long long cap = (long long)num_sms() * 16;
        if (grid > cap) grid = cap;
        if (smem > 48 * 1024)
            IMPDAR_CUDA(cudaFuncSetAttribute(winavg_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    ktimer_begin("winavg_kernel", (cudaStream_t)stream);
    winavg_kernel<T><<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(x, y, S, Tn, rows, half, taper, scratch);
    ktimer_end((cudaStream_t)stream);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

template <typename T>
static int rowabsmax_impl(const T *x, double *out, int S, int Tn, int batch, void *stream) {
    IMPDAR_CHECK_ARG(x && out, "rowabsmax: null pointer");
    IMPDAR_CHECK_ARG(S > 0 && Tn > 0 && batch > 0, "rowabsmax: bad shape");
    const long long rows = (long long)batch * S;
    long long grid = rows;
    const long long cap = (long long)num_sms() * 32;
    if (grid > cap) grid = cap;
    rowabsmax_kernel<T><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, out, Tn, rows);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

template <typename T>
static int rowgain_impl(const T *x, T *y, int S, int Tn, int batch, const double *gain, const int *trig, int in_double,
                        void *stream) {
    IMPDAR_CHECK_ARG(x && y && gain, "rowgain: null pointer");
    IMPDAR_CHECK_ARG(S > 0 && Tn > 0 && batch > 0, "rowgain: bad shape");
    const long long rows = (long long)batch * S;
    long long grid = rows;
    const long long cap = (long long)num_sms() * 32;
    if (grid > cap) grid = cap;
    rowgain_kernel<T><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, y, S, Tn, rows, gain, trig, in_double);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

}  // namespace impdar

using namespace impdar;

extern "C" {

size_t impdar_winavg_workspace_bytes(int S, int T, int batch) {
    (void)S;
    (void)batch;
    const size_t need_smem = (size_t)(T + 1) * sizeof(double);
    if (need_smem <= WINAVG_SMEM_LIMIT) return 0;
    return (size_t)num_sms() * 4 * (size_t)(T + 1) * sizeof(double);
}
int impdar_winavg_f32(const float *x, float *y, int S, int T, int batch, int half, const double *taper, void *ws,
                      size_t ws_bytes, void *stream) {
    return winavg_impl<float>(x, y, S, T, batch, half, taper, ws, ws_bytes, stream);
}
int impdar_winavg_f64(const double *x, double *y, int S, int T, int batch, int half, const double *taper, void *ws,
                      size_t ws_bytes, void *stream) {
    return winavg_impl<double>(x, y, S, T, batch, half, taper, ws, ws_bytes, stream);
}

size_t impdar_filtfilt_rows_workspace_bytes(int S, int T, int batch, int padlen, int elem_bytes) {
    return (size_t)batch * (size_t)S * (size_t)(T + 2 * padlen) * (size_t)elem_bytes;
}
int impdar_filtfilt_rows_f32(const float *x, float *y, int S, int T, int batch, const double *b, const double *a,
                             int ncoef, const double *zi, int padlen, void *ws, size_t ws_bytes, void *stream) {
    return filtfilt_rows_impl<float>(x, y, S, T, batch, b, a, ncoef, zi, padlen, ws, ws_bytes, stream);
}
int impdar_filtfilt_rows_f64(const double *x, double *y, int S, int T, int batch, const double *b, const double *a,
                             int ncoef, const double *zi, int padlen, void *ws, size_t ws_bytes, void *stream) {
    return filtfilt_rows_impl<double>(x, y, S, T, batch, b, a, ncoef, zi, padlen, ws, ws_bytes, stream);
}

int impdar_rowabsmax_f32(const float *x, double *out, int S, int T, int batch, void *stream) {
    return rowabsmax_impl<float>(x, out, S, T, batch, stream);
}
int impdar_rowabsmax_f64(const double *x, double *out, int S, int T, int batch, void *stream) {
    return rowabsmax_impl<double>(x, out, S, T, batch, stream);
}
int impdar_rowgain_f32(const float *x, float *y, int S, int T, int batch, const double *gain, const int *trig,
                       int in_double, void *stream) {
    return rowgain_impl<float>(x, y, S, T, batch, gain, trig, in_double, stream);
}
int impdar_rowgain_f64(const double *x, double *y, int S, int T, int batch, const double *gain, const int *trig,
                       int in_double, void *stream) {
    return rowgain_impl<double>(x, y, S, T, batch, gain, trig, in_double, stream);
}

}  // extern "C"
