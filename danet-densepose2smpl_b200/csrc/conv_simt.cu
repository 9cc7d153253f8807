// fp32 implicit-GEMM convolution on the FMA pipe (NHWC) -- the exact-parity path of the
// network half and the fallback for shapes the tcgen05 kernel does not take.
// Replaces nn.Conv2d (+ folded BatchNorm2d + ReLU + residual add) of models/module/hr_module.py,
// models/module/res_module.py; grouped convolutions (res_module.py:335-342,500-535) are expressed
// as `wsets` weight sets over the (batch,part)-flattened image axis (see include/danet_b200.h).
//
// CTA tile: 64 (image, output pixel) rows of one weight set x 64 output channels, 256 threads, 4x4 outputs per
// thread, K streamed in (tap, 16-channel) chunks through shared memory with register prefetch.
#include "common.cuh"

namespace danet {

constexpr int kTP = 64;      // pixels per CTA
constexpr int kTCo = 64;     // output channels per CTA
constexpr int kKC = 16;      // input channels per K chunk
constexpr int kAPitch = kTP + 4;

struct ConvArgs {
    int N, H, W, Cin, Cout, ks, stride, pad, Ho, Wo, wsets, relu;
    const float* x; const float* w; const float* bias; const float* res; float* y;
};

__global__ void __launch_bounds__(256)
k_conv_simt(ConvArgs a) {
    __shared__ __align__(16) float As[kKC][kAPitch];
    __shared__ __align__(16) float Bs[kKC][kTCo];
    // rows of the implicit GEMM = (image of this weight set, output pixel), flattened: image
    // n = g + wsets * (row / HoWo).  Tiny maps (4x4, 2x2) then still fill the 64-row tile.
    const int tid = threadIdx.x;
    const int g = blockIdx.z;
    const int q0 = blockIdx.x * kTP;
    const int co0 = blockIdx.y * kTCo;
    const int HoWo = a.Ho * a.Wo;
    const int rows = ((a.N - g + a.wsets - 1) / a.wsets) * HoWo;
    const int K = a.ks * a.ks * a.Cin;
    const float* wg = a.w + (size_t)g * K * a.Cout;

    // A-load role: row lp = tid/4, channel vec lv = tid%4
    const int lp = tid >> 2, lv = tid & 3;
    const int lq = q0 + lp;
    const bool lvalid = lq < rows;
    const int lk = lvalid ? lq / HoWo : 0, lpix = lvalid ? lq - lk * HoWo : 0;
    const int loh = lpix / a.Wo, low = lpix - loh * a.Wo;
    const float* xn = a.x + (size_t)(g + a.wsets * lk) * a.H * a.W * a.Cin;
    // B-load role: row bk = tid/16, col vec bv = tid%16
    const int bk = tid >> 4, bv = tid & 15;
    // compute role
    const int ty = tid >> 4, tx = tid & 15;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int cchunks = (a.Cin + kKC - 1) / kKC;
    const int nchunks = a.ks * a.ks * cchunks;

    auto load = [&](int chunk, float4& ra, float4& rb) {
        const int tap = chunk / cchunks, c0 = (chunk % cchunks) * kKC;
        const int r = tap / a.ks, s = tap % a.ks;
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        const int ih = loh * a.stride - a.pad + r, iw = low * a.stride - a.pad + s;
        const int c = c0 + lv * 4;
        if (lvalid && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W && c < a.Cin)
            ra = __ldg(reinterpret_cast<const float4*>(xn + ((size_t)ih * a.W + iw) * a.Cin + c));
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        const int kc = c0 + bk, co = co0 + bv * 4;
        if (kc < a.Cin && co < a.Cout)
            rb = __ldg(reinterpret_cast<const float4*>(wg + ((size_t)tap * a.Cin + kc) * a.Cout + co));
    };

    float4 ra, rb;
    load(0, ra, rb);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        As[lv * 4 + 0][lp] = ra.x; As[lv * 4 + 1][lp] = ra.y; As[lv * 4 + 2][lp] = ra.z; As[lv * 4 + 3][lp] = ra.w;
        *reinterpret_cast<float4*>(&Bs[bk][bv * 4]) = rb;
        __syncthreads();
        if (chunk + 1 < nchunks) load(chunk + 1, ra, rb);
#pragma unroll
        for (int k = 0; k < kKC; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float am[4] = {av.x, av.y, av.z, av.w};
            const float bm[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bm[j], acc[i][j]);
        }
        __syncthreads();
    }

    const int co = co0 + tx * 4;
    if (co >= a.Cout) return;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) bias = __ldg(reinterpret_cast<const float4*>(a.bias + (size_t)g * a.Cout + co));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ty * 4 + i;
        if (q >= rows) continue;
        const int qk = q / HoWo, qp = q - qk * HoWo;
        const size_t o = ((size_t)(g + a.wsets * qk) * HoWo + qp) * a.Cout + co;
        float4 v = make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
        if (a.res) {
            const float4 rr = __ldg(reinterpret_cast<const float4*>(a.res + o));
            v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
        }
        if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(a.y + o) = v;
    }
}

int conv_simt_launch(const danet_conv_desc* d, const float* x, const float* w, const float* bias,
                     const float* residual, float* y, cudaStream_t stream) {
    ConvArgs a;
    a.N = d->N; a.H = d->H; a.W = d->W; a.Cin = d->Cin; a.Cout = d->Cout; a.ks = d->ksize;
    a.stride = d->stride; a.pad = d->pad; a.wsets = d->wsets; a.relu = d->relu;
    a.Ho = (d->H + 2 * d->pad - d->ksize) / d->stride + 1;
    a.Wo = (d->W + 2 * d->pad - d->ksize) / d->stride + 1;
    a.x = x; a.w = w; a.bias = bias; a.res = residual; a.y = y;
    const int rows = cdiv(a.N, a.wsets) * a.Ho * a.Wo;
    dim3 grid(cdiv(rows, kTP), cdiv(a.Cout, kTCo), a.wsets);
    k_conv_simt<<<grid, 256, 0, stream>>>(a);
    DANET_LAUNCH_CHECK();
    return 0;
}

int conv_tc_launch(const danet_conv_desc* d, const float* x, const void* w_packed, const float* bias,
                   const float* residual, float* y, cudaStream_t stream);   // conv_tc.cu

}  // namespace danet

using namespace danet;

static int check_conv_desc(const danet_conv_desc* d) {
    DANET_CHECK(d, "danet_conv2d: null descriptor");
    DANET_CHECK(d->N >= 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "danet_conv2d: bad sizes");
    DANET_CHECK(d->Cin % 4 == 0 && d->Cout % 4 == 0, "danet_conv2d: Cin (%d) and Cout (%d) must be multiples of 4 (pad channels)", d->Cin, d->Cout);
    DANET_CHECK(d->ksize >= 1 && d->ksize <= 7 && d->stride >= 1 && d->stride <= 2 && d->pad >= 0, "danet_conv2d: bad ksize/stride/pad");
    DANET_CHECK(d->wsets >= 1, "danet_conv2d: wsets must be >= 1");
    DANET_CHECK(d->H + 2 * d->pad >= d->ksize && d->W + 2 * d->pad >= d->ksize, "danet_conv2d: kernel larger than padded input");
    DANET_CHECK(d->wsets <= 65535, "danet_conv2d: wsets=%d exceeds 65535", d->wsets);
    return 0;
}

extern "C" int danet_conv2d(const danet_conv_desc* d, int32_t algo, const float* x, const float* w,
                            const float* bias, const float* residual, float* y, danet_stream_t stream) {
    if (check_conv_desc(d) != 0) return -1;
    if (d->N == 0) return 0;
    DANET_CHECK(x && w && y, "danet_conv2d: null pointer");
    if (algo == DANET_CONV_SIMT) return conv_simt_launch(d, x, w, bias, residual, y, (cudaStream_t)stream);
    if (algo == DANET_CONV_TC) {
        DANET_CHECK(danet_conv_tc_supported(d), "danet_conv2d: shape not supported by the tcgen05 path");
        return conv_tc_launch(d, x, (const void*)w, bias, residual, y, (cudaStream_t)stream);
    }
    DANET_CHECK(false, "danet_conv2d: unknown algo %d", algo);
}
