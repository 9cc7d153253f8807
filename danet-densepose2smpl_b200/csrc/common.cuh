// Shared helpers for libdanet_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <vector>
#include "../../include/danet_b200.h"

namespace danet {

void set_error(const char* fmt, ...);

#define DANET_CHECK(cond, ...)                                   \
    do { if (!(cond)) { danet::set_error(__VA_ARGS__); return -1; } } while (0)

#define DANET_CUDA(expr)                                                              \
    do { cudaError_t _e = (expr);                                                     \
         if (_e != cudaSuccess) {                                                     \
             danet::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                    \
             return -2; } } while (0)

#define DANET_LAUNCH_CHECK()                                                          \
    do { cudaError_t _e = cudaGetLastError();                                         \
         if (_e != cudaSuccess) {                                                     \
             danet::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                    \
             return -3; } } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

template <typename T>
int upload(T** dptr, const T* host, size_t count) {
    DANET_CUDA(cudaMalloc((void**)dptr, count * sizeof(T) + 16));
    DANET_CUDA(cudaMemcpy(*dptr, host, count * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// cudaFuncSetAttribute is per device: true the first time the calling thread's current device is seen by this
// call site (`mask` is the site's static bit set, one bit per device ordinal; -1 on error)
inline int first_use_on_current_device(unsigned long long* mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    const bool first = !((*mask >> dev) & 1ull);
    *mask |= 1ull << dev;
    return first ? 1 : 0;
}

// two fp32 -> packed fp16x2 (lo in the lower half), round-to-nearest-even, saturating: the conversion the
// tensor-core convolution applies to its activations (conv_tc.cu), shared by the kernels that may write
// fp16 activation buffers for it
__device__ __forceinline__ uint32_t pack_h2_rn(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

}  // namespace danet
