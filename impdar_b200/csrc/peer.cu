// Peer-mapped images for the multi-GPU Kirchhoff migration (reference: the serial trace loop of
// migrationlib/mig_python.py:35-60 split into output-trace ranges, one process per GPU).
//
// The rank that holds the radargram owns one cudaMalloc'ed (snum, tnum) image; every other process of the node maps
// it into its own address space (CUDA IPC, peer access over NVLink / NVSwitch enabled lazily by the driver) and passes
// image + x_begin with ldo = tnum to impdar_kirchhoff_window_f32: the diffraction-sum kernels then store their output
// block straight into its final place in the holder's memory while they run - compute and "gather" are one kernel,
// there is no staging buffer, no collective and no copy after the last kernel.  Plain cudaMalloc memory (not a
// stream-ordered or virtual-memory pool) because that is what legacy IPC handles can export.
#include "common.cuh"

using namespace impdar;

extern "C" {

int impdar_peer_alloc(size_t bytes, void **ptr, void *handle64) {
    IMPDAR_CHECK_ARG(ptr && handle64 && bytes > 0, "peer_alloc: null pointer / zero size");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("peer_alloc: cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
        return IMPDAR_B200_ECUDA;
    }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        cudaGetLastError();
        set_error("peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        return IMPDAR_B200_ECUDA;
    }
    memcpy(handle64, &h, sizeof(h));
    *ptr = p;
    return IMPDAR_B200_OK;
}

int impdar_peer_free(void *ptr) {
    if (ptr) IMPDAR_CUDA(cudaFree(ptr));
    return IMPDAR_B200_OK;
}

int impdar_peer_open(const void *handle64, void **ptr) {
    IMPDAR_CHECK_ARG(ptr && handle64, "peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();   // not sticky: the caller falls back to the collective gather, later launches must not see it
        set_error("peer_open: cudaIpcOpenMemHandle -> %s", cudaGetErrorString(e));
        return IMPDAR_B200_ECUDA;
    }
    *ptr = p;
    return IMPDAR_B200_OK;
}

int impdar_peer_close(void *ptr) {
    if (ptr) IMPDAR_CUDA(cudaIpcCloseMemHandle(ptr));
    return IMPDAR_B200_OK;
}

int impdar_copy2d_f32(const float *src, size_t lds, float *dst, size_t ldd, int rows, int cols, void *stream) {
    IMPDAR_CHECK_ARG(src && dst && rows >= 0 && cols >= 0 && lds >= (size_t)cols && ldd >= (size_t)cols,
                     "copy2d: bad block %d x %d / strides", rows, cols);
    if (rows == 0 || cols == 0) return IMPDAR_B200_OK;
    IMPDAR_CUDA(cudaMemcpy2DAsync(dst, ldd * sizeof(float), src, lds * sizeof(float), (size_t)cols * sizeof(float),
                                  (size_t)rows, cudaMemcpyDefault, (cudaStream_t)stream));
    return IMPDAR_B200_OK;
}

}  // extern "C"
