// denoise (RadarData/_RadarDataFiltering.py:552-587): scipy.signal.wiener(data, mysize=(vert_win, hor_win), noise).
//   lMean = boxsum(x) / n,  lVar = boxsum(x^2) / n - lMean^2   (zero padded 'same' correlation with a ones window,
//                                                               window rows [s - V/2, s + (V-1)/2], columns likewise)
//   noise = mean(lVar) unless given;  out = lVar < noise ? lMean : (x - lMean) * (1 - noise / lVar) + lMean
// Sums and the variance are float64 (E[x^2] - mean^2 cancels catastrophically in float32); the output is float64 like
// scipy's.  Two launches: the variance pass reduces sum(lVar) and flags exact zeros (scipy divides by them: the reference
// turns that FloatingPointError into a ValueError), the apply pass recomputes the box sums - cheaper than parking two
// float64 images in HBM (8 B/sample read twice through L1/L2 instead of 4 + 16 + 16 + 8 B/sample).
#include "common.cuh"

namespace impdar {

template <typename T>
__device__ __forceinline__ void wiener_stats(const T *__restrict__ x, int S, int Tn, int s, int t, int V, int H, double n,
                                             double &mean, double &var) {
    const int r0 = max(0, s - V / 2), r1 = min(S - 1, s + (V - 1) / 2);
    const int c0 = max(0, t - H / 2), c1 = min(Tn - 1, t + (H - 1) / 2);
    double sum = 0.0, sq = 0.0;
    for (int r = r0; r <= r1; ++r) {
        const T *row = x + (size_t)r * Tn;
        for (int c = c0; c <= c1; ++c) {
            const T v = row[c];
            const T v2 = v * v;          // scipy squares in the input precision (im ** 2) before the float64 correlation
            sum += (double)v;
            sq += (double)v2;
        }
    }
    mean = sum / n;
    var = sq / n - mean * mean;
}

// acc[0] += sum of lVar, flags[0] |= 1 where lVar == 0, |= 2 where lVar is not finite
template <typename T>
__global__ void __launch_bounds__(256) wiener_var_kernel(const T *__restrict__ x, int S, int Tn, int V, int H,
                                                         double *__restrict__ acc, int *__restrict__ flags) {
    __shared__ double part[8];
    const double n = (double)V * (double)H;
    double local = 0.0;
    int f = 0;
    for (int s = blockIdx.y; s < S; s += gridDim.y)
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Tn; t += gridDim.x * blockDim.x) {
            double mean, var;
            wiener_stats(x, S, Tn, s, t, V, H, n, mean, var);
            local += var;
            if (var == 0.0) f |= 1;
        }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
    f = __syncthreads_or(f);
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < 8; ++i) tot += part[i];
        atomicAdd(acc, tot);
        if (f) atomicOr(flags, f);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) wiener_apply_kernel(const T *__restrict__ x, double *__restrict__ y, int S, int Tn,
                                                           int V, int H, const double *__restrict__ acc, double noise_in,
                                                           int estimate) {
    const double n = (double)V * (double)H;
    const double noise = estimate ? acc[0] / ((double)S * (double)Tn) : noise_in;
    for (int s = blockIdx.y; s < S; s += gridDim.y)
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Tn; t += gridDim.x * blockDim.x) {
            double mean, var;
            wiener_stats(x, S, Tn, s, t, V, H, n, mean, var);
            double res = (double)x[(size_t)s * Tn + t] - mean;
            res = __dmul_rn(res, 1.0 - noise / var);
            res = __dadd_rn(res, mean);
            y[(size_t)s * Tn + t] = (var < noise) ? mean : res;
        }
}

template <typename T>
static int wiener_impl(const T *x, double *y, int S, int Tn, int V, int H, int estimate, double noise, double *scratch,
                       void *stream) {
    IMPDAR_CHECK_ARG(x && y && scratch, "wiener: null pointer");
    IMPDAR_CHECK_ARG(S >= 1 && Tn >= 1 && V >= 1 && H >= 1, "wiener: bad shape or window");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)min((Tn + 255) / 256, 32), (unsigned)min(S, 8 * num_sms()));
    int *flags = reinterpret_cast<int *>(scratch + 1);
    IMPDAR_CUDA(cudaMemsetAsync(scratch, 0, 16, st));
    if (estimate) {
        ktimer_begin("wiener_var_kernel", st);
        wiener_var_kernel<T><<<grid, 256, 0, st>>>(x, S, Tn, V, H, scratch, flags);
        ktimer_end(st);
        IMPDAR_LAUNCH_CHECK();
    }
    ktimer_begin("wiener_apply_kernel", st);
    wiener_apply_kernel<T><<<grid, 256, 0, st>>>(x, y, S, Tn, V, H, scratch, noise, estimate);
    ktimer_end(st);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

// ---------------------------------------------------------------------------------------------- median filter
// denoise(ftype='median') = scipy.ndimage.median_filter(data, size=(V, H)) (mode 'reflect', origin 0): the element of
// rank (V*H)//2 of the window rows [s - V/2, s - V/2 + V), columns [t - H/2, t - H/2 + H), indices reflected about the
// edges with the edge sample repeated (d c b a | a b c d | d c b a).  Pure selection - bit-exact.  The window is copied
// once into the thread's local array and the wanted rank is found by counting (n^2 / 2 comparisons, no data movement).
constexpr int MEDIAN_MAX_WINDOW = 256;

__device__ __forceinline__ int reflect_index(int i, int n) {
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}

template <typename T>
__global__ void __launch_bounds__(128) median_kernel(const T *__restrict__ x, T *__restrict__ y, int S, int Tn, int V, int H) {
    T win[MEDIAN_MAX_WINDOW];
    const int n = V * H, want = n / 2;
    for (int s = blockIdx.y; s < S; s += gridDim.y)
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Tn; t += gridDim.x * blockDim.x) {
            int k = 0;
            for (int dv = 0; dv < V; ++dv) {
                const T *row = x + (size_t)reflect_index(s - V / 2 + dv, S) * Tn;
                for (int dh = 0; dh < H; ++dh) win[k++] = row[reflect_index(t - H / 2 + dh, Tn)];
            }
            T out = win[0];
            for (int i = 0; i < n; ++i) {
                const T v = win[i];
                int less = 0, equal_before = 0;
                for (int j = 0; j < n; ++j) {
                    less += (win[j] < v);
                    equal_before += (j < i) & (win[j] == v);
                }
                if (less + equal_before == want) {      // stable rank of element i
                    out = v;
                    break;
                }
            }
            y[(size_t)s * Tn + t] = out;
        }
}

template <typename T>
static int median_impl(const T *x, T *y, int S, int Tn, int V, int H, void *stream) {
    IMPDAR_CHECK_ARG(x && y, "median: null pointer");
    IMPDAR_CHECK_ARG(S >= 1 && Tn >= 1 && V >= 1 && H >= 1, "median: bad shape or window");
    IMPDAR_CHECK_ARG((long long)V * H <= MEDIAN_MAX_WINDOW, "median: window of %d x %d samples exceeds the %d supported", V, H,
                     MEDIAN_MAX_WINDOW);
    IMPDAR_CHECK_ARG((const void *)x != (const void *)y, "median: in-place not supported");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)min((Tn + 127) / 128, 64), (unsigned)min(S, 16 * num_sms()));
    ktimer_begin("median_kernel", st);
    median_kernel<T><<<grid, 128, 0, st>>>(x, y, S, Tn, V, H);
    ktimer_end(st);
    IMPDAR_LAUNCH_CHECK();
    return IMPDAR_B200_OK;
}

}  // namespace impdar

using namespace impdar;

extern "C" {

int impdar_wiener_f32(const float *x, double *y, int snum, int tnum, int vert_win, int hor_win, int estimate_noise,
                      double noise, double *scratch, void *stream) {
    return wiener_impl<float>(x, y, snum, tnum, vert_win, hor_win, estimate_noise, noise, scratch, stream);
}
int impdar_wiener_f64(const double *x, double *y, int snum, int tnum, int vert_win, int hor_win, int estimate_noise,
                      double noise, double *scratch, void *stream) {
    return wiener_impl<double>(x, y, snum, tnum, vert_win, hor_win, estimate_noise, noise, scratch, stream);
}

int impdar_median_f32(const float *x, float *y, int snum, int tnum, int vert_win, int hor_win, void *stream) {
    return median_impl<float>(x, y, snum, tnum, vert_win, hor_win, stream);
}
int impdar_median_f64(const double *x, double *y, int snum, int tnum, int vert_win, int hor_win, void *stream) {
    return median_impl<double>(x, y, snum, tnum, vert_win, hor_win, stream);
}

}  // extern "C"
