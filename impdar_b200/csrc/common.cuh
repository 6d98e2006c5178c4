// Shared helpers for libimpdar_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/impdar_b200.h"

namespace impdar {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// Optional CUDA-event bracket around one kernel launch (impdar_b200_kernel_timer, for bench.py's roofline:
// the dominant kernel's own duration on the stream it is launched on).  No-ops unless enabled.
void ktimer_begin(const char *kernel, cudaStream_t st);
void ktimer_end(cudaStream_t st);

#define IMPDAR_CHECK_ARG(cond, ...)            \
    do {                                       \
        if (!(cond)) {                         \
            impdar::set_error(__VA_ARGS__);    \
            return IMPDAR_B200_EINVAL;         \
        }                                      \
    } while (0)

#define IMPDAR_CUDA(call)                                                                    \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            impdar::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return IMPDAR_B200_ECUDA;                                                        \
        }                                                                                    \
    } while (0)

#define IMPDAR_LAUNCH_CHECK()                                                                \
    do {                                                                                     \
        impdar::count_launch();                                                              \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess) {                                                            \
            impdar::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return IMPDAR_B200_ECUDA;                                                        \
        }                                                                                    \
    } while (0)

// Function attributes (dynamic shared-memory limits) are per device: cache "already set" per device index, not per process.
constexpr int IMPDAR_MAX_DEVICES = 64;
static inline int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < IMPDAR_MAX_DEVICES) ? dev : 0;
}

static inline int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit accesses that do not pollute L1
__device__ __forceinline__ float4 ld_stream4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w));
}

}  // namespace impdar
