// TMA box-load throughput probe (bring-up tool, not part of the product): how fast does one SM / the whole chip pull
// conv halo boxes [C x W x H] of an NHWC fp16 tensor through cp.async.bulk.tensor.4d, and 1-D bulk copies of the same size?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu ; run: ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// mode 0: tensor-map box loads; mode 1: 1-D bulk copies of `bytes` from a linear buffer
__global__ void k_probe(const __grid_constant__ CUtensorMap tm, const uint8_t* lin, int mode, int bytes, int iters, int N, int H, int W,
                        int bw, int bh, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int stages = 3;
    const uint32_t bars = base + stages * ((bytes + 1023) / 1024 * 1024);
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        int img = blockIdx.x % N, th = 0, tw = 0;
        for (int i = 0; i < iters + stages; ++i) {
            const int s = i % stages;
            if (i >= stages) mbar_wait(bars + 8 * s, ((i / stages) - 1) & 1);
            if (i < iters) {
                const uint32_t dst = base + s * ((bytes + 1023) / 1024 * 1024);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8 * s), "r"((uint32_t)bytes) : "memory");
                if (mode == 0) {
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                                 ::"r"(dst), "l"(&tm), "r"(0), "r"(tw * (bw - 2) - 1), "r"(th * (bh - 2) - 1), "r"(img), "r"(bars + 8 * s) : "memory");
                    tw += 1; if (tw * (bw - 2) >= W) { tw = 0; th += 1; if (th * (bh - 2) >= H) { th = 0; img = (img + 1) % N; } }
                } else {
                    const size_t off = ((size_t)(blockIdx.x * 131 + i) * 4096) % ((size_t)N * H * W * 96);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(lin + (off & ~15ull)), "r"((uint32_t)bytes), "r"(bars + 8 * s) : "memory");
                }
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled)p;
    const int N = 64, H = 56, W = 56;
    long long* d_cyc; cudaMalloc(&d_cyc, 148 * 8);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Case { int C, kch, bw, bh, swz; const char* name; };
    Case cases[] = {{48, 64, 18, 18, 128, "C48 box 64x18x18 sw128"}, {48, 64, 10, 18, 128, "C48 box 64x10x18 sw128"},
                    {64, 64, 18, 18, 128, "C64 box 64x18x18 sw128"}, {96, 64, 18, 18, 128, "C96 box 64x18x18 sw128"},
                    {192, 64, 18, 18, 128, "C192 box 64x18x18 sw128"}, {24, 32, 16, 16, 64, "C24 box 32x16x16 sw64"},
                    {48, 16, 18, 18, 32, "C48 box 16x18x18 sw32"}};
    for (auto& c : cases) {
        __half* x; size_t n = (size_t)N * H * W * c.C;
        cudaMalloc(&x, n * 2 + 4096); cudaMemset(x, 0, n * 2);
        CUtensorMap tm;
        cuuint64_t gdim[4] = {(cuuint64_t)c.C, W, H, N};
        cuuint64_t gstr[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)W * c.C * 2, (cuuint64_t)H * W * c.C * 2};
        cuuint32_t box[4] = {(cuuint32_t)c.kch, (cuuint32_t)c.bw, (cuuint32_t)c.bh, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUtensorMapSwizzle sw = c.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (c.swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
        const int bytes = c.kch * 2 * c.bw * c.bh;
        for (int grid : {1, 148}) for (int mode : {0, 1}) {
            const int iters = 400;
            k_probe<<<grid, 32, 3 * ((bytes + 1023) / 1024 * 1024) + 2048>>>(tm, (const uint8_t*)x, mode, bytes, iters, N, H, W, c.bw, c.bh, d_cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
            std::vector<long long> cyc(148); cudaMemcpy(cyc.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
            printf("%-28s grid %3d %s: %6.0f cycles/box (%5d B) = %5.1f B/clk/SM, rows %d -> %.1f cycles/row\n", c.name, grid, mode ? "bulk-1D" : "tensor ",
                   (double)mx / iters, bytes, (double)bytes * iters / mx, c.bw * c.bh, (double)mx / iters / (c.bw * c.bh));
        }
        cudaFree(x);
    }
    return 0;
}
