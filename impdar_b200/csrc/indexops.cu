// Index / resampling operations either side of the hot path, so that a radargram can stay
// in HBM across a whole processing chain.  Reference: RadarData/_RadarDataProcessing.py
//   :20-47   reverse           (np.fliplr)
//   :238-349 crop              (row block copy; 'pretrig' with a trigger vector = per-trace shift with NaN fill)
//   :352-421 hcrop             (column block copy)
//   :424-477 restack           (np.mean over groups of `traces` columns - numpy's pairwise summation order)
//   :66-188  nmo               (per-trace scipy interp1d(kind='linear') onto a new time vector = row interpolation)
//   :50-63   constant_sample_depth_spacing (row interpolation)
//   :499-584 constant_space    (column compaction + column interpolation)
//   :587-637 elev_correct      (per-trace downward shift with NaN fill)
// Everything here is bit-exact against the reference (numpy 2.3 / scipy 1.18 arithmetic order): no FMA contraction
// in the interpolation formulas, IEEE division, numpy's pairwise summation tree for the means.
// All kernels are pure HBM streams: lanes run along the trace axis (contiguous), 8 B/sample of compulsory traffic.
#include "common.cuh"

namespace impdar {

constexpr int IO_ROWS = 4;   // rows per thread in the streaming kernels below

static inline dim3 io_grid(int ncols, int nrows, int batch = 1) {
    const int gx = min((ncols + 255) / 256, 64);
    const int row_groups = (nrows + IO_ROWS - 1) / IO_ROWS;
    int gy = (8 * num_sms() + gx - 1) / gx;          // ~8 CTAs of 256 threads per SM over the whole grid
    if (gy > row_groups) gy = row_groups;
    if (gy < 1) gy = 1;
    return dim3((unsigned)gx, (unsigned)gy, (unsigned)batch);
}

// ------------------------------------------------------------------------------------------------ block copy
// y[b, i, j] = x[b, r0 + i, flip ? c0 + (nc - 1 - j) : c0 + j]
template <typename T>
__global__ void __launch_bounds__(256) crop_block_kernel(const T *__restrict__ x, T *__restrict__ y, int S, int T_,
                                                         int nr, int nc, int r0, int c0, int flip) {
    const int b = blockIdx.z;
    const T *xb = x + (size_t)b * S * T_;
    T *yb = y + (size_t)b * nr * nc;
    // IO_ROWS rows per thread: independent loads in flight (a 4-byte element per thread and row would leave the
    // memory system idle: ~8 KB in flight per SM against the ~40 KB the HBM latency needs)
    for (int i0 = blockIdx.y * IO_ROWS; i0 < nr; i0 += gridDim.y * IO_ROWS) {
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nc; j += gridDim.x * blockDim.x) {
            const int js = flip ? (nc - 1 - j) : j;
            T v[IO_ROWS];
#pragma unroll
            for (int k = 0; k < IO_ROWS; ++k)
                if (i0 + k < nr) v[k] = xb[(size_t)(r0 + i0 + k) * T_ + c0 + js];
#pragma unroll
            for (int k = 0; k < IO_ROWS; ++k)
                if (i0 + k < nr) yb[(size_t)(i0 + k) * nc + j] = v[k];
        }
    }
}

template <typename T>
static int crop_block(const T *x, T *y, int S, int T_, int batch, int r0, int r1, int c0, int c1, int flip,
                      void *stream) {
    IMPDAR_CHECK_ARG(x && y, "crop: null pointer");
    IMPDAR_CHECK_ARG(S >= 1 && T_ >= 1 && batch >= 1, "crop: bad shape");
    IMPDAR_CHECK_ARG(0 <= r0 && r0 <= r1 && r1 <= S && 0 <= c0 && c0 <= c1 && c1 <= T_, "crop: bad limits");
    const int nr = r1 - r0, nc = c1 - c0;
    if (nr == 0 || nc == 0) return IMPDAR_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid = io_grid(nc, nr, batch);
    ktimer_begin("crop_block_kernel", st);
    crop_block_kernel<T><<<grid, 256, 0, st>>>(x, y, S, T_, nr, nc, r0, c0, flip);
    ktimer_end(st);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

// --------------------------------------------------------------------------------- per-trace shift, NaN fill
// y[i, t] = x[i + shift[t], t] where 0 <= i + shift[t] < S_in, NaN elsewhere.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) shift_traces_kernel(const TI *__restrict__ x, TO *__restrict__ y, int S_in,
                                                           int T_, int S_out, const int *__restrict__ shift) {
    const TO nanv = (TO)__longlong_as_double(0x7ff8000000000000LL);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T_; t += gridDim.x * blockDim.x) {
        const int sh = shift[t];
        for (int i0 = blockIdx.y * IO_ROWS; i0 < S_out; i0 += gridDim.y * IO_ROWS) {
            TO v[IO_ROWS];
#pragma unroll
            for (int k = 0; k < IO_ROWS; ++k) {
                const long long src = (long long)i0 + k + sh;
                v[k] = (i0 + k < S_out && src >= 0 && src < S_in) ? (TO)x[(size_t)src * T_ + t] : nanv;
            }
#pragma unroll
            for (int k = 0; k < IO_ROWS; ++k)
                if (i0 + k < S_out) y[(size_t)(i0 + k) * T_ + t] = v[k];
        }
    }
}

template <typename TI, typename TO>
static int shift_traces(const TI *x, TO *y, int S_in, int T_, int S_out, const int *shift, void *stream) {
    IMPDAR_CHECK_ARG(x && y && shift, "shift_traces: null pointer");
    IMPDAR_CHECK_ARG(S_in >= 1 && T_ >= 1 && S_out >= 0, "shift_traces: bad shape");
    if (S_out == 0) return IMPDAR_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid = io_grid(T_, S_out);
    ktimer_begin("shift_traces_kernel", st);
    shift_traces_kernel<TI, TO><<<grid, 256, 0, st>>>(x, y, S_in, T_, S_out, shift);
    ktimer_end(st);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- restack
// numpy's add.reduce over a contiguous run: 0 + pairwise_sum(a, n)  (numpy/_core/src/umath/loops_utils.h.src):
//   n < 8: sequential;  n <= 128: eight running sums over blocks of 8, combined as a balanced tree, remainder
//   sequential;  n > 128: split at (n/2 rounded down to a multiple of 8), recurse.
template <typename T>
__device__ __forceinline__ T np_pairwise_leaf(const T *a, int n) {   // n <= 128
    if (n < 8) {
        T r = (T)0;
        for (int i = 0; i < n; ++i) r = r + a[i];
        return r;
    }
    T r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 = r0 + a[i + 0];
        r1 = r1 + a[i + 1];
        r2 = r2 + a[i + 2];
        r3 = r3 + a[i + 3];
        r4 = r4 + a[i + 4];
        r5 = r5 + a[i + 5];
        r6 = r6 + a[i + 6];
        r7 = r7 + a[i + 7];
    }
    T res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res = res + a[i];
    return res;
}

template <typename T>
__device__ T np_pairwise_sum(const T *a, int n) {
    if (n <= 128) return np_pairwise_leaf(a, n);
    // explicit stack instead of recursion: segments are visited left to right, partial sums combined in the
    // same (left + right) tree numpy's recursion builds.  Depth <= 32 for any int n.
    struct Frame {
        int off, n, state;
        T left;
    };
    Frame stk[32];
    int sp = 0;
    stk[0] = {0, n, 0, (T)0};
    T ret = (T)0;
    while (sp >= 0) {
        Frame &f = stk[sp];
        if (f.n <= 128) {
            ret = np_pairwise_leaf(a + f.off, f.n);
            --sp;
            continue;
        }
        int n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.state == 0) {
            f.state = 1;
            stk[sp + 1] = {f.off, n2, 0, (T)0};
            ++sp;
        } else if (f.state == 1) {
            f.left = ret;
            f.state = 2;
            stk[sp + 1] = {f.off + n2, f.n - n2, 0, (T)0};
            ++sp;
        } else {
            ret = f.left + ret;
            --sp;
        }
    }
    return ret;
}

// y[s, j] = (TO)( (0 + pairwise_sum(x[s, j*n : (j+1)*n])) / n ), the sum and the division in the input precision
// (np.mean keeps float32 for float32 input; _RadarDataProcessing.py:456 stores into a float64 array).
template <typename TI, typename TO>
__global__ void __launch_bounds__(128) restack_kernel(const TI *__restrict__ x, TO *__restrict__ y, int S, int T_,
                                                      int To, int n) {
    const TI cnt = (TI)n;
    for (int s = blockIdx.y; s < S; s += gridDim.y) {
        const TI *row = x + (size_t)s * T_;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < To; j += gridDim.x * blockDim.x) {
            const TI sum = (TI)0 + np_pairwise_sum(row + (size_t)j * n, n);
            y[(size_t)s * To + j] = (TO)(sum / cnt);
        }
    }
}

template <typename TI, typename TO>
static int restack(const TI *x, TO *y, int S, int T_, int n, void *stream) {
    IMPDAR_CHECK_ARG(x && y, "restack: null pointer");
    IMPDAR_CHECK_ARG(S >= 1 && T_ >= 1 && n >= 1, "restack: bad shape");
    const int To = T_ / n;
    if (To == 0) return IMPDAR_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)min((To + 127) / 128, 32), (unsigned)min(S, 8 * num_sms()));
    ktimer_begin("restack_kernel", st);
    restack_kernel<TI, TO><<<grid, 128, 0, st>>>(x, y, S, T_, To, n);
    ktimer_end(st);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

// ---------------------------------------------------------------------------------------- linear interpolation
// Two arithmetic forms, both used by the reference through scipy.interpolate.interp1d(kind='linear'):
//  mode 0  scipy's own _call_linear (any 2-D y, or float32 y):
//            y = w_hi * y[hi] + w_lo * y[lo],   w_hi = (x-x_lo)/(x_hi-x_lo),  w_lo = (x_hi-x)/(x_hi-x_lo)
//  mode 1  numpy.interp (1-D float64 y; what interp1d dispatches to inside nmo's per-trace loop):
//            slope = (y[j+1]-y[j])/(xp[j+1]-xp[j]);  y = slope*(x-xp[j]) + y[j]
//            exact hit x == xp[j] -> y[j];  NaN -> retry from the right node; still NaN and y[j]==y[j+1] -> y[j]
// The node weights are O(n) float64 host work and arrive as a table of `InterpNode`.
struct InterpNode {
    int lo, hi;     // source rows (or columns)
    double a, b;    // mode 0: w_hi, w_lo.  mode 1: x - xp[lo], x - xp[hi]
    double den;     // mode 1: xp[hi] - xp[lo]
    int exact;      // mode 1: 1 = copy y[lo]
    int pad_;
};

template <typename TI>
__device__ __forceinline__ double interp_value(const InterpNode &nd, TI ylo, TI yhi, int mode) {
    if (mode == 0) return __dadd_rn(__dmul_rn(nd.a, (double)yhi), __dmul_rn(nd.b, (double)ylo));
    if (nd.exact) return (double)ylo;
    const double dlo = (double)ylo, dhi = (double)yhi;
    const double slope = __ddiv_rn(__dsub_rn(dhi, dlo), nd.den);
    double r = __dadd_rn(__dmul_rn(slope, nd.a), dlo);
    if (isnan(r)) {
        r = __dadd_rn(__dmul_rn(slope, nd.b), dhi);
        if (isnan(r) && dlo == dhi) r = dlo;
    }
    return r;
}

// rows: y[i, t] = f(x[lo_i, t], x[hi_i, t])
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) interp_rows_kernel(const TI *__restrict__ x, TO *__restrict__ y, int T_,
                                                          int S_out, const InterpNode *__restrict__ nodes, int mode) {
    for (int i0 = blockIdx.y * IO_ROWS; i0 < S_out; i0 += gridDim.y * IO_ROWS) {
        InterpNode nd[IO_ROWS];
#pragma unroll
        for (int k = 0; k < IO_ROWS; ++k) nd[k] = nodes[min(i0 + k, S_out - 1)];
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T_; t += gridDim.x * blockDim.x) {
            TI lo[IO_ROWS], hi[IO_ROWS];
#pragma unroll
            for (int k = 0; k < IO_ROWS; ++k) {
                lo[k] = x[(size_t)nd[k].lo * T_ + t];
                hi[k] = x[(size_t)nd[k].hi * T_ + t];
            }
#pragma unroll
            for (int k = 0; k < IO_ROWS; ++k)
                if (i0 + k < S_out) y[(size_t)(i0 + k) * T_ + t] = (TO)interp_value<TI>(nd[k], lo[k], hi[k], mode);
        }
    }
}

// columns: y[s, j] = f(x[s, lo_j], x[s, hi_j])
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) interp_cols_kernel(const TI *__restrict__ x, TO *__restrict__ y, int S,
                                                          int T_in, int T_out, const InterpNode *__restrict__ nodes,
                                                          int mode) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < T_out; j += gridDim.x * blockDim.x) {
        const InterpNode nd = nodes[j];
        for (int s0 = blockIdx.y * IO_ROWS; s0 < S; s0 += gridDim.y * IO_ROWS) {
            TI lo[IO_ROWS], hi[IO_ROWS];
#pragma unroll
            for (int k = 0; k < IO_ROWS; ++k) {
                const TI *row = x + (size_t)min(s0 + k, S - 1) * T_in;
                lo[k] = row[nd.lo];
                hi[k] = row[nd.hi];
            }
#pragma unroll
            for (int k = 0; k < IO_ROWS; ++k)
                if (s0 + k < S) y[(size_t)(s0 + k) * T_out + j] = (TO)interp_value<TI>(nd, lo[k], hi[k], mode);
        }
    }
}

template <typename TI, typename TO>
static int interp_rows(const TI *x, TO *y, int S_in, int T_, int S_out, const void *nodes, int mode, void *stream) {
    IMPDAR_CHECK_ARG(x && y && nodes, "interp_rows: null pointer");
    IMPDAR_CHECK_ARG(S_in >= 2 && T_ >= 1 && S_out >= 0 && (mode == 0 || mode == 1), "interp_rows: bad argument");
    if (S_out == 0) return IMPDAR_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid = io_grid(T_, S_out);
    ktimer_begin("interp_rows_kernel", st);
    interp_rows_kernel<TI, TO><<<grid, 256, 0, st>>>(x, y, T_, S_out, (const InterpNode *)nodes, mode);
    ktimer_end(st);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

template <typename TI, typename TO>
static int interp_cols(const TI *x, TO *y, int S, int T_in, int T_out, const void *nodes, int mode, void *stream) {
    IMPDAR_CHECK_ARG(x && y && nodes, "interp_cols: null pointer");
    IMPDAR_CHECK_ARG(S >= 1 && T_in >= 2 && T_out >= 0 && (mode == 0 || mode == 1), "interp_cols: bad argument");
    if (T_out == 0) return IMPDAR_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid = io_grid(T_out, S);
    ktimer_begin("interp_cols_kernel", st);
    interp_cols_kernel<TI, TO><<<grid, 256, 0, st>>>(x, y, S, T_in, T_out, (const InterpNode *)nodes, mode);
    ktimer_end(st);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

}  // namespace impdar

using namespace impdar;

extern "C" {

size_t impdar_interp_node_bytes(void) { return sizeof(InterpNode); }

int impdar_crop_f32(const float *x, float *y, int snum, int tnum, int batch, int r0, int r1, int c0, int c1,
                    int flip_lr, void *stream) {
    return impdar_crop_bytes(x, y, snum, tnum, batch, r0, r1, c0, c1, flip_lr, 4, stream);
}
int impdar_crop_f64(const double *x, double *y, int snum, int tnum, int batch, int r0, int r1, int c0, int c1,
                    int flip_lr, void *stream) {
    return impdar_crop_bytes(x, y, snum, tnum, batch, r0, r1, c0, c1, flip_lr, 8, stream);
}

int impdar_crop_bytes(const void *x, void *y, int snum, int tnum, int batch, int r0, int r1, int c0, int c1,
                      int flip_lr, int elem_bytes, void *stream) {
    // 128-bit lanes whenever every row segment starts and ends on a 16-byte boundary
    if (!flip_lr && elem_bytes >= 1 && elem_bytes < 16 && 16 % elem_bytes == 0) {
        const int per = 16 / elem_bytes;
        if (tnum % per == 0 && c0 % per == 0 && c1 % per == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0)
            return crop_block<uint4>((const uint4 *)x, (uint4 *)y, snum, tnum / per, batch, r0, r1, c0 / per, c1 / per, 0, stream);
    }
    switch (elem_bytes) {
        case 1: return crop_block<uint8_t>((const uint8_t *)x, (uint8_t *)y, snum, tnum, batch, r0, r1, c0, c1, flip_lr, stream);
        case 2: return crop_block<uint16_t>((const uint16_t *)x, (uint16_t *)y, snum, tnum, batch, r0, r1, c0, c1, flip_lr, stream);
        case 4: return crop_block<uint32_t>((const uint32_t *)x, (uint32_t *)y, snum, tnum, batch, r0, r1, c0, c1, flip_lr, stream);
        case 8: return crop_block<uint64_t>((const uint64_t *)x, (uint64_t *)y, snum, tnum, batch, r0, r1, c0, c1, flip_lr, stream);
        case 16: return crop_block<uint4>((const uint4 *)x, (uint4 *)y, snum, tnum, batch, r0, r1, c0, c1, flip_lr, stream);
        default: break;
    }
    set_error("crop: element size %d not in {1, 2, 4, 8, 16}", elem_bytes);
    return IMPDAR_B200_EINVAL;
}

int impdar_shift_traces_f32(const float *x, float *y, int snum_in, int tnum, int snum_out, const int *shift,
                            void *stream) {
    return shift_traces<float, float>(x, y, snum_in, tnum, snum_out, shift, stream);
}
int impdar_shift_traces_f32_f64(const float *x, double *y, int snum_in, int tnum, int snum_out, const int *shift,
                                void *stream) {
    return shift_traces<float, double>(x, y, snum_in, tnum, snum_out, shift, stream);
}
int impdar_shift_traces_f64(const double *x, double *y, int snum_in, int tnum, int snum_out, const int *shift,
                            void *stream) {
    return shift_traces<double, double>(x, y, snum_in, tnum, snum_out, shift, stream);
}

int impdar_restack_f32(const float *x, float *y, int snum, int tnum, int traces, void *stream) {
    return restack<float, float>(x, y, snum, tnum, traces, stream);
}
int impdar_restack_f32_f64(const float *x, double *y, int snum, int tnum, int traces, void *stream) {
    return restack<float, double>(x, y, snum, tnum, traces, stream);
}
int impdar_restack_f64(const double *x, double *y, int snum, int tnum, int traces, void *stream) {
    return restack<double, double>(x, y, snum, tnum, traces, stream);
}

int impdar_interp_rows_f32(const float *x, float *y, int snum_in, int tnum, int snum_out, const void *nodes,
                           int mode, void *stream) {
    return interp_rows<float, float>(x, y, snum_in, tnum, snum_out, nodes, mode, stream);
}
int impdar_interp_rows_f32_f64(const float *x, double *y, int snum_in, int tnum, int snum_out, const void *nodes,
                               int mode, void *stream) {
    return interp_rows<float, double>(x, y, snum_in, tnum, snum_out, nodes, mode, stream);
}
int impdar_interp_rows_f64(const double *x, double *y, int snum_in, int tnum, int snum_out, const void *nodes,
                           int mode, void *stream) {
    return interp_rows<double, double>(x, y, snum_in, tnum, snum_out, nodes, mode, stream);
}

int impdar_interp_cols_f32(const float *x, float *y, int snum, int tnum_in, int tnum_out, const void *nodes,
                           int mode, void *stream) {
    return interp_cols<float, float>(x, y, snum, tnum_in, tnum_out, nodes, mode, stream);
}
int impdar_interp_cols_f32_f64(const float *x, double *y, int snum, int tnum_in, int tnum_out, const void *nodes,
                               int mode, void *stream) {
    return interp_cols<float, double>(x, y, snum, tnum_in, tnum_out, nodes, mode, stream);
}
int impdar_interp_cols_f64(const double *x, double *y, int snum, int tnum_in, int tnum_out, const void *nodes,
                           int mode, void *stream) {
    return interp_cols<double, double>(x, y, snum, tnum_in, tnum_out, nodes, mode, stream);
}

}  // extern "C"
