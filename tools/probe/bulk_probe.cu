// Micro-probe: rate of small 1-D cp.async.bulk copies (shared<->global) issued by many threads of a CTA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_probe bulk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256) k_store(float* out, int bytes, int pitch, int iters, long long* cyc) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < 256 * pitch / 4; i += 256) ((float*)sm)[i] = i;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const long long t0 = clock64();
    uint8_t* g = (uint8_t*)out + ((size_t)blockIdx.x * 256 + threadIdx.x) * 1024;
    for (int it = 0; it < iters; ++it) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + (it & 3) * 256), "r"(s32(sm + threadIdx.x * pitch)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if ((it & 3) == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
}
__global__ void __launch_bounds__(256) k_load(const float* in, int bytes, int pitch, int iters, long long* cyc, float* sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(256)); }
    __syncthreads();
    const long long t0 = clock64();
    const uint8_t* g = (const uint8_t*)in + ((size_t)blockIdx.x * 256 + threadIdx.x) * 1024;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(sm + threadIdx.x * pitch)), "l"(g + (it & 3) * 256), "r"(bytes), "r"(s32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(&bar)), "r"(ph) : "memory");
        ph ^= 1;
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = clock64() - t0; sink[0] = ((float*)sm)[3]; }
}
int main() {
    float* buf; long long* cyc; cudaMalloc(&buf, (size_t)148 * 256 * 1024 + 4096); cudaMalloc(&cyc, 64);
    cudaMemset(buf, 0, (size_t)148 * 256 * 1024);
    cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 64;
    for (int grid : {1, 148})
        for (int bytes : {64, 96, 128, 192, 256}) {
            const int pitch = bytes + 16;
            long long c;
            k_store<<<grid, 256, 256 * pitch>>>(buf, bytes, pitch, iters, cyc); k_store<<<grid, 256, 256 * pitch>>>(buf, bytes, pitch, iters, cyc);
            cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("store grid=%3d bytes=%3d : %7.1f cycles per 256 copies (%.1f cyc/copy, %.1f B/cyc/SM) err=%s\n", grid, bytes, (double)c / iters, (double)c / iters / 256, 256.0 * bytes * iters / c, cudaGetErrorString(cudaGetLastError()));
            k_load<<<grid, 256, 256 * pitch>>>(buf, bytes, pitch, iters, cyc, buf); k_load<<<grid, 256, 256 * pitch>>>(buf, bytes, pitch, iters, cyc, buf);
            cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("load  grid=%3d bytes=%3d : %7.1f cycles per 256 copies incl. round trip (%.1f B/cyc/SM) err=%s\n", grid, bytes, (double)c / iters, 256.0 * bytes * iters / c, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
