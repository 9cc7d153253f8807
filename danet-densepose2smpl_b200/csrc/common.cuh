// Shared helpers for libdanet_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <mutex>
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
inline std::mutex& first_use_mutex() { static std::mutex m; return m; }
inline int first_use_on_current_device(unsigned long long* mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    std::lock_guard<std::mutex> lk(first_use_mutex());       // host threads may make their first launch concurrently
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

// Device view of a danet_act (include/danet_b200.h): fp32 tensor and/or split-fp16 planes of one NHWC activation.
// Readers prefer the fp32 view when present; writers fill every view that is present.  The split is the one the
// tensor-core convolution's epilogue applies: hi = rn_f16(v) (saturating), lo = rn_f16(v - hi).
struct ActV { float* f; __half* hi; __half* lo; };
static inline ActV actv(const danet_act* a) {
    ActV v; v.f = a ? a->f32 : nullptr; v.hi = a ? (__half*)a->hi : nullptr; v.lo = a ? (__half*)a->lo : nullptr;
    return v;
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t u) {
    float2 r;
    asm("{\n\t.reg .f16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}" : "=f"(r.x), "=f"(r.y) : "r"(u));
    return r;
}
// four consecutive channels starting at element e (e % 4 == 0; planes are 8-byte aligned there)
__device__ __forceinline__ float4 act_ld4(const ActV& a, size_t e) {
    if (a.f) return __ldg(reinterpret_cast<const float4*>(a.f + e));
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(a.hi + e));
    float2 p = h2_to_f2(h.x), q = h2_to_f2(h.y);
    if (a.lo) {
        const uint2 l = __ldg(reinterpret_cast<const uint2*>(a.lo + e));
        const float2 pl = h2_to_f2(l.x), ql = h2_to_f2(l.y);
        p.x += pl.x; p.y += pl.y; q.x += ql.x; q.y += ql.y;
    }
    return make_float4(p.x, p.y, q.x, q.y);
}
__device__ __forceinline__ void act_st4(const ActV& a, size_t e, float4 v) {
    if (a.f) *reinterpret_cast<float4*>(a.f + e) = v;
    if (a.hi) {
        const uint32_t h0 = pack_h2_rn(v.x, v.y), h1 = pack_h2_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(a.hi + e) = make_uint2(h0, h1);
        if (a.lo) {
            const float2 p = h2_to_f2(h0), q = h2_to_f2(h1);
            *reinterpret_cast<uint2*>(a.lo + e) = make_uint2(pack_h2_rn(v.x - p.x, v.y - p.y), pack_h2_rn(v.z - q.x, v.w - q.y));
        }
    }
}
// eight consecutive channels starting at element e (e % 8 == 0): 16-byte accesses on the fp16 planes
struct float8 { float4 a, b; };
__device__ __forceinline__ float8 act_ld8(const ActV& a, size_t e) {
    float8 r;
    if (a.f) { r.a = __ldg(reinterpret_cast<const float4*>(a.f + e)); r.b = __ldg(reinterpret_cast<const float4*>(a.f + e + 4)); return r; }
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(a.hi + e));
    float2 p0 = h2_to_f2(h.x), p1 = h2_to_f2(h.y), p2 = h2_to_f2(h.z), p3 = h2_to_f2(h.w);
    if (a.lo) {
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(a.lo + e));
        const float2 q0 = h2_to_f2(l.x), q1 = h2_to_f2(l.y), q2 = h2_to_f2(l.z), q3 = h2_to_f2(l.w);
        p0.x += q0.x; p0.y += q0.y; p1.x += q1.x; p1.y += q1.y; p2.x += q2.x; p2.y += q2.y; p3.x += q3.x; p3.y += q3.y;
    }
    r.a = make_float4(p0.x, p0.y, p1.x, p1.y); r.b = make_float4(p2.x, p2.y, p3.x, p3.y);
    return r;
}
__device__ __forceinline__ void act_st8(const ActV& a, size_t e, const float8& v) {
    if (a.f) { *reinterpret_cast<float4*>(a.f + e) = v.a; *reinterpret_cast<float4*>(a.f + e + 4) = v.b; }
    if (a.hi) {
        uint4 h;
        h.x = pack_h2_rn(v.a.x, v.a.y); h.y = pack_h2_rn(v.a.z, v.a.w); h.z = pack_h2_rn(v.b.x, v.b.y); h.w = pack_h2_rn(v.b.z, v.b.w);
        *reinterpret_cast<uint4*>(a.hi + e) = h;
        if (a.lo) {
            const float2 p0 = h2_to_f2(h.x), p1 = h2_to_f2(h.y), p2 = h2_to_f2(h.z), p3 = h2_to_f2(h.w);
            uint4 l;
            l.x = pack_h2_rn(v.a.x - p0.x, v.a.y - p0.y); l.y = pack_h2_rn(v.a.z - p1.x, v.a.w - p1.y);
            l.z = pack_h2_rn(v.b.x - p2.x, v.b.y - p2.y); l.w = pack_h2_rn(v.b.z - p3.x, v.b.w - p3.y);
            *reinterpret_cast<uint4*>(a.lo + e) = l;
        }
    }
}
__device__ __forceinline__ float act_ld1(const ActV& a, size_t e) {
    if (a.f) return __ldg(a.f + e);
    float v = __half2float(a.hi[e]);
    if (a.lo) v += __half2float(a.lo[e]);
    return v;
}
__device__ __forceinline__ void act_st1(const ActV& a, size_t e, float v) {
    if (a.f) a.f[e] = v;
    if (a.hi) {
        const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
        a.hi[e] = h;
        if (a.lo) a.lo[e] = __float2half_rn(v - __half2float(h));
    }
}
static inline bool act_any(const danet_act* a) { return a && (a->f32 || a->hi); }

}  // namespace danet
