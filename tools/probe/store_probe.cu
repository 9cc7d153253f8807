// Micro-probe: cost of the epilogue's global-store patterns (8 warps per CTA, one CTA per SM, 32 KB "tiles").
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe store_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// tile = 128 pixels x 64 fp32 channels (256 B per pixel), pixels of a tile are 8-pixel runs of a 56-wide image row
template <int P>
__global__ void __launch_bounds__(256) k(float* out, int tiles, long long* cyc, size_t img_stride) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64();
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        float* base = out + (size_t)t * 128 * 64;               // tile-contiguous is enough for the probe
        const float v = (float)t;
        if (P == 1) {            // lane = row, 4 x STG.128 per 16-column group; warp w: rows 32*(w&3).., groups (w>>2), +2
            const int row = (warp & 3) * 32 + lane;
            for (int g = warp >> 2; g < 4; g += 2)
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(base + row * 64 + g * 16 + 4 * j) = make_float4(v, v, v, v);
        } else if (P == 2) {     // quad: 8 rows x 32 B per STG.64
            const int r0 = (warp & 3) * 32 + (lane >> 2), c = 2 * (lane & 3);
            for (int g = warp >> 2; g < 4; g += 2)
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) *reinterpret_cast<float2*>(base + (r0 + 8 * k) * 64 + g * 16 + 8 * i + c) = make_float2(v, v);
        } else if (P == 3) {     // quad after exchange: 8 rows x 64 B per STG.128
            const int r0 = (warp & 3) * 32 + (lane >> 2), c = 4 * (lane & 3);
            for (int g = warp >> 2; g < 4; g += 2)
#pragma unroll
                for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(base + (r0 + 8 * k) * 64 + g * 16 + c) = make_float4(v, v, v, v);
        } else {                 // fully coalesced: 4 lines x 128 B per STG.128 (2 rows per instruction)
            for (int it = warp; it < 64; it += 8) *reinterpret_cast<float4*>(base + it * 128 + lane * 4) = make_float4(v, v, v, v);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
}
int main() {
    const int tiles = 148 * 290;
    float* buf; long long* cyc;
    cudaMalloc(&buf, (size_t)tiles * 128 * 64 * 4); cudaMalloc(&cyc, 64);
    for (int rep = 0; rep < 2; ++rep)
        for (int p = 1; p <= 4; ++p) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            if (p == 1) k<1><<<148, 256>>>(buf, tiles, cyc, 0);
            if (p == 2) k<2><<<148, 256>>>(buf, tiles, cyc, 0);
            if (p == 3) k<3><<<148, 256>>>(buf, tiles, cyc, 0);
            if (p == 4) k<4><<<148, 256>>>(buf, tiles, cyc, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("pattern %d: %.1f us, %.0f cycles per 32 KB tile, %.2f TB/s  (%s)\n", p, ms * 1e3, (double)c / 290.0,
                   (double)tiles * 32768 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
