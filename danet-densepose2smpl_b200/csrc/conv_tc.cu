// tcgen05 implicit-GEMM convolution engine for sm_100a, second generation.
// The tensor-core path of models/module/hr_module.py:161-179,334-378 + res_module.py:27-97,281-390,393-535
// convolutions (conv + folded BN + residual + ReLU): 1x1 / 3x3 / 7x7, stride 1 or 2, weight sets (the
// reference's grouped convolutions over the (batch, part)-flattened image axis).
//
// Numerics.  Activations and weights are SPLIT-FP16: v = hi + lo with hi = rn_f16(v), lo = rn_f16(v - hi)
// (22 significant bits, fp16 exponent range; plain fp32 numbers below 6e-5 keep an absolute error <= 3e-8).
// "exact" mode issues three MMAs per K step, hi*hi + hi*lo + lo*hi, into one fp32 TMEM accumulator
// (the dropped lo*lo term is 2^-22 relative): fp32-grade results from the fp16 tensor pipe, which is what
// lets the default path meet the reference's fp32 outputs to 1e-4.  "fast" mode issues hi*hi only.
//
// Structure (one persistent CTA per SM, warp-specialised, up to kMaxProb independent convolutions -- e.g. the
// parallel branches of one HRNet stage -- in ONE launch over a concatenated tile space):
//   A (activations): fp16 NHWC planes in HBM.  One TMA tensor-map load (cp.async.bulk.tensor.4d, SWIZZLE_128B/64B/32B,
//       out-of-bounds zero fill = the convolution's padding, elementStrides = 2 for the parity planes of a
//       stride-2 convolution) brings the input HALO of a tile -- (16+k-1) x (8*S+k-1) pixels x <= 64 channels --
//       into shared memory, one 128/64/32-byte row per pixel.  Eight consecutive MMA rows are eight consecutive
//       pixels of one halo row, the next 8-row group is the next image row (SBO = halo pitch), so every filter
//       tap is just a different descriptor start address: an input element crosses L2->SM ~1.3x, not 9x.
//       S = 1 or 2 sub-tiles of 16 x 8 output pixels share one halo (256 pixels per pipeline step).
//   B (weights): pre-packed once (danet_conv_tc_pack) into the swizzled shared-memory image of every
//       (N tile, channel chunk, parity plane, tap group[, hi/lo]) block; streamed with 1-D cp.async.bulk.
//   MMA: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16); accumulators in TMEM,
//       two 256-column halves used alternately (a tile that needs more than 256 columns takes both).
//       exact mode with 2*N <= 256: the hi and lo weight rows are concatenated along N, so hi*[hi|lo] is ONE
//       MMA of width 2N (the A operand is fetched once) and the epilogue adds the two column ranges.
//   Epilogue: 8 warps, tcgen05.ld.16x256b (a quad of lanes owns one 32-byte sector of a pixel), bias + residual
//       + ReLU, output as split-fp16 planes and/or fp32.
//   Launch: programmatic dependent launch; every mbarrier wait is bounded (traps instead of hanging).
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <mutex>
#include <string.h>

namespace danet {
namespace tc {

constexpr int kTileH = 16, kTileW = 8;
constexpr int kEpiWarps = 8;
constexpr int kWarpA = 8, kWarpB = 9, kWarpMma = 10;   // the MMA issuer is the highest warp id (issue priority)
constexpr int kThreads = 11 * 32;
constexpr int kNumEpi = kEpiWarps * 32;
constexpr int kMaxAStages = 8, kMaxBStages = 8;
constexpr int kSmemMax = 227 * 1024;                   // opt-in dynamic shared memory per CTA on sm_100
constexpr int kSmemFixed = 2048;                       // barriers + 1024-byte alignment slack

struct alignas(64) Prob {
    CUtensorMap tm[2];                   // input planes: hi, lo
    const uint8_t* wpk; const float* bias;
    const float* res_f; const __half* res_hi; const __half* res_lo;
    float* y_f; __half* y_hi; __half* y_lo;
    int N, H, W, Cin, Cout, Ho, Wo, ks, stride, pad, relu, wsets, exact;
    int S;                               // sub-tiles (16 x 8 output pixels each) per pipeline step
    int npa;                             // active parity planes (1 for stride 1, up to 4 for stride 2)
    int SWB, KCH, nchunks;               // swizzle bytes per pixel row, channels per chunk, chunks
    int TG;                              // filter taps per weight block
    int NT, ntn;                         // output channels per N tile, N tiles
    int nconcat;                         // exact mode: hi and lo weight rows share one block (one MMA of width 2*NT)
    int wsplit;                          // exact mode without nconcat: hi and lo rows are separate blocks
    int ACC;                             // accumulator columns per sub-tile
    int big;                             // S*ACC > 256: the tile takes both accumulator halves
    int par_py[4], par_px[4], ntap[4], ngrp[4], stage_bytes[4], sbo_a[4];
    int tapoff16[4][16];                 // smem offset (16-byte units) of each tap's shifted view inside the plane
    int tapidx[4][16];                   // original filter tap index r*ks+s (weight packing)
    int tiles_w, tiles_h, tile_count, tile_base;
    int rows_blk;                        // rows of one weight block per tap
    int tap_bytes, b_block_bytes, nblk;  // nblk: weight blocks per (weight set, N tile)
    int bpc;                             // weight blocks per channel chunk
    long long blocks_per_set;
    unsigned long long m_ntn, m_tw, m_th, m_ws;
};

constexpr int kMaxProb = 6;
struct ArgsN {
    int nprob, total_tiles, na_stages, a_slot_bytes, nb_stages, b_slot_bytes;
    long long* prof;
    Prob p[kMaxProb];
};

static unsigned long long magic40(int d) { return (1ull << 40) / (unsigned long long)d + 1ull; }
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

// ---------------------------------------------------------------------------------------------
// host: geometry of one problem (everything but the shared-memory ring sizes, which belong to the launch)
// ---------------------------------------------------------------------------------------------
static bool make_prob(const danet_conv_desc* d, int S_req, Prob* g) {
    if (!(d->stride == 1 || d->stride == 2) || !(d->ksize == 1 || d->ksize == 3 || d->ksize == 7) || d->pad != d->ksize / 2) return false;
    if (d->Cin % 8 != 0 || d->Cout % 8 != 0 || d->H < 1 || d->W < 1 || d->N < 1 || d->wsets < 1) return false;
    g->N = d->N; g->H = d->H; g->W = d->W; g->Cin = d->Cin; g->Cout = d->Cout; g->ks = d->ksize;
    g->pad = d->pad; g->stride = d->stride; g->relu = d->relu; g->wsets = d->wsets;
    g->exact = (d->flags & DANET_CONV_EXACT) ? 1 : 0;
    g->Ho = (d->H + 2 * d->pad - d->ksize) / d->stride + 1;
    g->Wo = (d->W + 2 * d->pad - d->ksize) / d->stride + 1;
    if (g->Ho < 1 || g->Wo < 1) return false;
    // N tiling / accumulators
    const int np = (d->Cout + 15) / 16 * 16;
    g->ntn = (np + 255) / 256;
    g->NT = ((np + g->ntn - 1) / g->ntn + 15) / 16 * 16;
    g->nconcat = (g->exact && 2 * g->NT <= 256 && env_int("DANET_TC_NCONCAT", 1)) ? 1 : 0;
    g->wsplit = (g->exact && !g->nconcat) ? 1 : 0;
    g->ACC = g->NT * (g->nconcat ? 2 : 1);
    int S = S_req;
    if (g->Wo <= kTileW) S = 1;
    if (S * g->ACC > 512) S = 1;
    g->S = S;
    g->big = S * g->ACC > 256 ? 1 : 0;
    // parity decomposition: a stride-2 convolution is the sum over the input parities (py,px) of dense
    // stride-1 sub-convolutions; each parity plane is one TMA box with elementStrides = 2
    g->npa = 0;
    int max_ntap = 0, tr_max = 0, tc_max = 0;
    int tc_[4];
    for (int py = 0; py < d->stride; ++py)
        for (int px = 0; px < d->stride; ++px) {
            const int tr = py < d->ksize ? (d->ksize - 1 - py) / d->stride + 1 : 0;     // taps r = py + stride*i
            const int tcn = px < d->ksize ? (d->ksize - 1 - px) / d->stride + 1 : 0;
            if (tr * tcn == 0) continue;
            if (tr * tcn > 16) return false;
            const int a = g->npa++;
            g->par_py[a] = py; g->par_px[a] = px; g->ntap[a] = tr * tcn; tc_[a] = tcn;
            max_ntap = tr * tcn > max_ntap ? tr * tcn : max_ntap;
            tr_max = tr > tr_max ? tr : tr_max; tc_max = tcn > tc_max ? tcn : tc_max;
        }
    // every parity plane is loaded with the box of the largest one (one tensor map per input plane)
    const int Hb = kTileH + tr_max - 1, Wb = kTileW * S + tc_max - 1;
    // swizzle width: the widest row unless the channel count is tiny
    int swb = 128;
    const int c16 = (d->Cin + 15) / 16 * 16;
    while (swb > 32 && swb / 2 >= 2 * c16) swb /= 2;
    swb = env_int("DANET_TC_SWB", swb);
    g->SWB = swb; g->KCH = swb / 2;
    g->nchunks = (d->Cin + g->KCH - 1) / g->KCH;
    g->rows_blk = g->NT * (g->nconcat ? 2 : 1);
    g->tap_bytes = g->rows_blk * swb;
    int tg = max_ntap;
    while (tg > 1 && tg * g->tap_bytes > 24 * 1024) --tg;
    g->TG = tg;
    g->b_block_bytes = (tg * g->tap_bytes + 1023) / 1024 * 1024;
    g->bpc = 0;
    for (int a = 0; a < 4; ++a) {
        if (a >= g->npa) { g->par_py[a] = g->par_px[a] = g->ntap[a] = g->ngrp[a] = g->stage_bytes[a] = g->sbo_a[a] = 0;
                           for (int k = 0; k < 16; ++k) g->tapoff16[a][k] = g->tapidx[a][k] = 0; continue; }
        g->stage_bytes[a] = Hb * Wb * swb;
        g->sbo_a[a] = Wb * swb;
        g->ngrp[a] = (g->ntap[a] + tg - 1) / tg;
        g->bpc += g->ngrp[a] * (g->wsplit ? 2 : 1);
        for (int k = 0; k < 16; ++k) { g->tapoff16[a][k] = 0; g->tapidx[a][k] = 0; }
        for (int k = 0; k < g->ntap[a]; ++k) {
            const int ti = k / tc_[a], tj = k % tc_[a];
            g->tapoff16[a][k] = ((ti * Wb + tj) * swb) >> 4;
            g->tapidx[a][k] = (g->par_py[a] + d->stride * ti) * d->ksize + (g->par_px[a] + d->stride * tj);
        }
    }
    g->nblk = g->nchunks * g->bpc;
    g->blocks_per_set = (long long)g->ntn * g->nblk;
    g->tiles_w = (g->Wo + kTileW * S - 1) / (kTileW * S); g->tiles_h = (g->Ho + kTileH - 1) / kTileH;
    const long long tiles = (long long)d->N * g->tiles_h * g->tiles_w * g->ntn;
    if (tiles >= (1 << 24) || g->wsets >= (1 << 16)) return false;
    if ((long long)d->N * g->Ho * g->Wo * d->Cout >= (1LL << 31) || (long long)d->N * d->H * d->W * d->Cin >= (1LL << 31)) return false;   // 32-bit element offsets
    g->tile_count = (int)tiles; g->tile_base = 0;
    g->m_ntn = magic40(g->ntn); g->m_tw = magic40(g->tiles_w); g->m_th = magic40(g->tiles_h); g->m_ws = magic40(g->wsets);
    return true;
}
static int max_stage_bytes(const Prob& g) { int m = 0; for (int a = 0; a < g.npa; ++a) m = g.stage_bytes[a] > m ? g.stage_bytes[a] : m; return m; }

// ring sizes of a launch over n problems; false if they cannot fit even with S = 1 everywhere
static bool plan_rings(ArgsN* a) {
    int amax = 0, bmax = 0, need_a = 2;
    for (int i = 0; i < a->nprob; ++i) {
        const int sb = max_stage_bytes(a->p[i]);
        amax = sb > amax ? sb : amax;
        bmax = a->p[i].b_block_bytes > bmax ? a->p[i].b_block_bytes : bmax;
        if (a->p[i].exact) need_a = 4;
    }
    a->a_slot_bytes = (amax + 1023) / 1024 * 1024;
    a->b_slot_bytes = (bmax + 1023) / 1024 * 1024;
    int nb = 3;
    int na = (kSmemMax - kSmemFixed - nb * a->b_slot_bytes) / a->a_slot_bytes;
    if (na < need_a) { nb = 2; na = (kSmemMax - kSmemFixed - nb * a->b_slot_bytes) / a->a_slot_bytes; }
    if (na < (need_a == 4 ? 2 : 2)) return false;
    if (na < need_a) return false;
    if (na > kMaxAStages) na = kMaxAStages;
    // spend what is left on deeper weight prefetch
    while (nb < kMaxBStages && kSmemFixed + na * a->a_slot_bytes + (nb + 1) * a->b_slot_bytes <= kSmemMax) ++nb;
    a->na_stages = na; a->nb_stages = nb;
    return true;
}

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();                                        // bounded wait: never hang the device
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// programmatic dependent launch: this grid may start while the previous kernel of the stream drains; nothing the
// previous kernel wrote (activations, residual) or still reads (our output may be its input) is touched before pdl_wait()
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// 16 lanes x 256 bits, twice (columns +0..7 and +8..15): thread t receives, for i = 0/1,
// r[4i], r[4i+1] = row t/4, columns 8i + 2*(t%4) + {0,1};  r[4i+2], r[4i+3] = row t/4 + 8, same columns
// (the mma m16n8 accumulator fragment).  A quad then owns 32 contiguous bytes of a row.
__device__ __forceinline__ void tc_ld16x256_x2_nowait(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1).
// layout_type: SWIZZLE_128B = 2, SWIZZLE_64B = 4, SWIZZLE_32B = 6; LBO is unused for swizzled K-major.
// The hardware derives the swizzle phase from the absolute shared-memory address, so a start address that is
// only row-aligned (a filter tap shifted by a few pixels) needs no base_offset (verified on B200, round 1).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
    return ((uint64_t)hi << 32) | lo;
}
// physical offset of logical byte offset `off` inside a 1024-byte-aligned swizzled region
__host__ __device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t mask) { return off ^ (((off >> 7) & mask) << 4); }
__device__ __forceinline__ int mdiv(int x, unsigned long long m) { return (int)(((unsigned long long)(unsigned)x * m) >> 40); }

struct TileCoord { int nt, tw, th, img; };
__device__ __forceinline__ TileCoord decode_tile(const Prob& g, int t) {
    TileCoord c;
    int r = mdiv(t, g.m_ntn); c.nt = t - r * g.ntn;
    int r2 = mdiv(r, g.m_tw); c.tw = r - r2 * g.tiles_w;
    c.img = mdiv(r2, g.m_th); c.th = r2 - c.img * g.tiles_h;
    return c;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
k_conv_tc(const __grid_constant__ ArgsN a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;          // swizzle atoms need 1024-byte alignment
    const uint32_t sA = sbase;
    const uint32_t sB = sbase + a.na_stages * a.a_slot_bytes;
    const uint32_t sBar = sB + a.nb_stages * a.b_slot_bytes;
    // barrier map (8 bytes each)
    const uint32_t bar_a_full = sBar, bar_a_empty = sBar + 8 * kMaxAStages;
    const uint32_t bar_b_full = sBar + 16 * kMaxAStages, bar_b_empty = bar_b_full + 8 * kMaxBStages;
    const uint32_t bar_acc_full = bar_b_empty + 8 * kMaxBStages, bar_acc_empty = bar_acc_full + 16;
    const uint32_t tmem_slot_addr = bar_acc_empty + 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar_acc_full + 8 * i, 1); mbar_init(bar_acc_empty + 8 * i, kNumEpi); }
        for (int i = 0; i < a.na_stages; ++i) { mbar_init(bar_a_full + 8 * i, 1); mbar_init(bar_a_empty + 8 * i, 1); }
        for (int i = 0; i < a.nb_stages; ++i) { mbar_init(bar_b_full + 8 * i, 1); mbar_init(bar_b_empty + 8 * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpA && lane < a.nprob) { tma_prefetch_desc(&a.p[lane].tm[0]); if (a.p[lane].exact) tma_prefetch_desc(&a.p[lane].tm[1]); }
    if (warp == kWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot_addr), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot_addr));
    pdl_launch_dependents();            // the next launch may fill SMs as our CTAs retire

    if (warp == kWarpA) {
        // ================= A producer: one TMA box per (plane, channel chunk, parity plane) =================
        if (lane == 0) {
            int as = 0; uint32_t aph = 0;
            int pi = 0;
            pdl_wait();                                              // activations come from the previous kernel
            for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
                while (tile >= a.p[pi].tile_base + a.p[pi].tile_count) ++pi;
                const Prob& P = a.p[pi];
                const TileCoord tc = decode_tile(P, tile - P.tile_base);
                const int h0 = tc.th * kTileH * P.stride - P.pad, w0 = tc.tw * kTileW * P.S * P.stride - P.pad;
                for (int c = 0; c < P.nchunks; ++c)
                    for (int slot = 0; slot < P.npa; ++slot)
                        for (int pl = 0; pl <= P.exact; ++pl) {
                            mbar_wait(bar_a_empty + 8 * as, ((aph >> as) & 1u) ^ 1u);
                            mbar_expect_tx(bar_a_full + 8 * as, (uint32_t)P.stage_bytes[slot]);
                            tma_load_4d(sA + as * a.a_slot_bytes, &P.tm[pl], c * P.KCH, w0 + P.par_px[slot], h0 + P.par_py[slot],
                                        tc.img, bar_a_full + 8 * as);
                            aph ^= 1u << as;
                            if (++as == a.na_stages) as = 0;
                        }
            }
        }
    } else if (warp == kWarpB) {
        // ================= B producer: bulk copies of pre-packed weight blocks =================
        if (lane == 0) {
            int bs = 0; uint32_t bph = 0;
            int pi = 0;
            for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
                while (tile >= a.p[pi].tile_base + a.p[pi].tile_count) ++pi;
                const Prob& P = a.p[pi];
                const TileCoord tc = decode_tile(P, tile - P.tile_base);
                const int ws = tc.img - mdiv(tc.img, P.m_ws) * P.wsets;
                const uint8_t* src = P.wpk + ((long long)ws * P.blocks_per_set + (long long)tc.nt * P.nblk) * P.b_block_bytes;
                for (int b = 0; b < P.nblk; ++b) {
                    mbar_wait(bar_b_empty + 8 * bs, ((bph >> bs) & 1u) ^ 1u);
                    mbar_expect_tx(bar_b_full + 8 * bs, (uint32_t)P.b_block_bytes);
                    bulk_g2s(sB + bs * a.b_slot_bytes, src + (long long)b * P.b_block_bytes, (uint32_t)P.b_block_bytes, bar_b_full + 8 * bs);
                    bph ^= 1u << bs;
                    if (++bs == a.nb_stages) bs = 0;
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ================= MMA issuer =================
        // The whole warp runs the loop (warp-uniform control flow keeps descriptors in uniform registers); one
        // elected lane issues the tcgen05 instructions.
        int as = 0, bs = 0; uint32_t aph = 0, bph = 0, eph = 0; int tog = 0;
        int pi = 0;
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
            while (tile >= a.p[pi].tile_base + a.p[pi].tile_count) ++pi;
            const Prob& P = a.p[pi];
            int cs = 0;
            if (P.big) {
                mbar_wait(bar_acc_empty, (eph & 1u) ^ 1u); mbar_wait(bar_acc_empty + 8, ((eph >> 1) & 1u) ^ 1u);
                eph ^= 3u;
            } else {
                cs = tog; tog ^= 1;
                mbar_wait(bar_acc_empty + 8 * cs, ((eph >> cs) & 1u) ^ 1u);
                eph ^= 1u << cs;
            }
            tc_fence_after();
            const uint32_t d_base = tmem_base + cs * 256;
            // kind::f16, A/B = F16 (format 0), D = F32, both K-major, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc1 = (1u << 4) | ((uint32_t)(P.NT >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * P.NT) >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t ltype = P.SWB == 128 ? 2u : (P.SWB == 64 ? 4u : 6u);
            const uint64_t bd0 = make_desc(0, 8 * P.SWB, ltype);
            const int kmma = P.KCH / 16;                         // K = 16 halves (32 bytes) per MMA
            const uint32_t tap16 = P.tap_bytes >> 4;
            const int mode = !P.exact ? 0 : (P.nconcat ? 1 : 2);
            const bool S2 = P.S == 2;
            const uint32_t ACC = (uint32_t)P.ACC;
            const uint32_t sub16 = (uint32_t)(kTileW * P.SWB) >> 4;
            uint32_t acc = 0;
            for (int c = 0; c < P.nchunks; ++c) {
                const int kreal = (P.Cin - c * P.KCH + 15) >> 4;
                const int kv = kreal < kmma ? kreal : kmma;      // K steps wholly beyond Cin are not issued
                for (int slot = 0; slot < P.npa; ++slot) {
                    const int as_hi = as;
                    mbar_wait(bar_a_full + 8 * as, (aph >> as) & 1u);
                    aph ^= 1u << as; if (++as == a.na_stages) as = 0;
                    int as_lo = as_hi;
                    if (P.exact) {
                        as_lo = as;
                        mbar_wait(bar_a_full + 8 * as, (aph >> as) & 1u);
                        aph ^= 1u << as; if (++as == a.na_stages) as = 0;
                    }
                    tc_fence_after();
                    const uint64_t ad0 = make_desc(0, (uint32_t)P.sbo_a[slot], ltype);
                    const uint64_t ad_hi = ad0 + ((sA + as_hi * a.a_slot_bytes) >> 4);
                    const uint64_t ad_lo = ad0 + ((sA + as_lo * a.a_slot_bytes) >> 4);
                    for (int tg = 0; tg < P.ngrp[slot]; ++tg) {
                        const int k0 = tg * P.TG;
                        const int ntk = min(P.TG, P.ntap[slot] - k0);
                        mbar_wait(bar_b_full + 8 * bs, (bph >> bs) & 1u);
                        tc_fence_after();
                        const uint64_t bd = bd0 + ((sB + bs * a.b_slot_bytes) >> 4);
                        if (elect_one()) {
                            for (int tt = 0; tt < ntk; ++tt) {
                                const uint32_t toff = (uint32_t)P.tapoff16[slot][k0 + tt];
                                for (int kk = 0; kk < kv; ++kk) {
                                    const uint64_t bdk = bd + tt * tap16 + 2 * kk;
                                    const uint64_t adh = ad_hi + toff + 2 * kk, adl = ad_lo + toff + 2 * kk;
                                    if (mode == 0) {
                                        tc_mma_f16(d_base, adh, bdk, idesc1, acc);
                                        if (S2) tc_mma_f16(d_base + ACC, adh + sub16, bdk, idesc1, acc);
                                    } else if (mode == 1) {                       // nconcat: hi * [hi | lo], then lo * hi
                                        tc_mma_f16(d_base, adh, bdk, idesc2, acc);
                                        if (S2) tc_mma_f16(d_base + ACC, adh + sub16, bdk, idesc2, acc);
                                        tc_mma_f16(d_base, adl, bdk, idesc1, 1u);
                                        if (S2) tc_mma_f16(d_base + ACC, adl + sub16, bdk, idesc1, 1u);
                                    } else {                                      // hi * hi, lo * hi (hi * lo from the next block)
                                        tc_mma_f16(d_base, adh, bdk, idesc1, acc);
                                        if (S2) tc_mma_f16(d_base + ACC, adh + sub16, bdk, idesc1, acc);
                                        tc_mma_f16(d_base, adl, bdk, idesc1, 1u);
                                        if (S2) tc_mma_f16(d_base + ACC, adl + sub16, bdk, idesc1, 1u);
                                    }
                                    acc = 1;
                                }
                            }
                            tc_commit(bar_b_empty + 8 * bs);
                        }
                        __syncwarp();
                        acc = 1;
                        bph ^= 1u << bs; if (++bs == a.nb_stages) bs = 0;
                        if (P.wsplit) {
                            // the lo weight rows of the same taps: hi * lo
                            mbar_wait(bar_b_full + 8 * bs, (bph >> bs) & 1u);
                            tc_fence_after();
                            const uint64_t bl = bd0 + ((sB + bs * a.b_slot_bytes) >> 4);
                            if (elect_one()) {
                                for (int tt = 0; tt < ntk; ++tt) {
                                    const uint32_t toff = (uint32_t)P.tapoff16[slot][k0 + tt];
                                    for (int kk = 0; kk < kv; ++kk) {
                                        tc_mma_f16(d_base, ad_hi + toff + 2 * kk, bl + tt * tap16 + 2 * kk, idesc1, 1u);
                                        if (S2) tc_mma_f16(d_base + ACC, ad_hi + toff + sub16 + 2 * kk, bl + tt * tap16 + 2 * kk, idesc1, 1u);
                                    }
                                }
                                tc_commit(bar_b_empty + 8 * bs);
                            }
                            __syncwarp();
                            bph ^= 1u << bs; if (++bs == a.nb_stages) bs = 0;
                        }
                    }
                    if (elect_one()) {
                        tc_commit(bar_a_empty + 8 * as_hi);
                        if (P.exact) tc_commit(bar_a_empty + 8 * as_lo);
                    }
                    __syncwarp();
                }
            }
            if (elect_one()) tc_commit(bar_acc_full + 8 * cs);
            __syncwarp();
        }
    } else {
        // ================= epilogue: TMEM -> bias/residual/ReLU -> global =================
        // Quad mapping (tcgen05.ld 16x256b): a thread owns tile column wwq = lane/4 of the four tile rows 4q..4q+3 and,
        // per 16-column group, channels 2*(lane%4)+{0,1} and +8: the four lanes of a quad read/write one whole 32-byte
        // sector; a warp instruction touches 8 lines instead of 32.  Each of the two warps of a TMEM lane quarter takes
        // every other 16-column group.  Residual + bias operands of up to three units ahead are held in registers and
        // the first three are requested BEFORE the accumulator wait.
        const int q = warp & 3;                                  // TMEM lane quarter this warp may access
        const int half = warp >> 2;                              // 0/1: which 16-column groups this warp owns
        const int wwq = lane >> 2, cq = 2 * (lane & 3);
        const bool odd = lane & 1;
        uint32_t fph = 0; int tog = 0;
        int pi = 0;
        pdl_wait();                                              // residual reads / output writes
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
            while (tile >= a.p[pi].tile_base + a.p[pi].tile_count) ++pi;
            const Prob& P = a.p[pi];
            const TileCoord tc = decode_tile(P, tile - P.tile_base);
            int cs = 0;
            if (!P.big) { cs = tog; tog ^= 1; }
            const int ngroups = P.NT / 16;
            const int gph = (ngroups - half + 1) >> 1;            // groups this warp owns per sub-tile
            const int nunits = P.S * gph;
            const int oh0 = tc.th * kTileH + 4 * q;
            const uint32_t rowstep = (uint32_t)(P.Wo * P.Cout);
            const int cw = P.Cout - tc.nt * P.NT;                 // channels of this N tile that exist (multiple of 8)
            const int boff = (tc.img - mdiv(tc.img, P.m_ws) * P.wsets) * P.Cout + tc.nt * P.NT + cq;
            const bool has_res = P.res_f != nullptr || P.res_hi != nullptr;
            // unit u -> (sub-tile s, column group grp); element offset of this thread's first pixel / channel
            auto unit_s = [&](int u) { return u / gph; };
            auto unit_grp = [&](int u) { return half + 2 * (u - (u / gph) * gph); };
            auto pix_of = [&](int s) {
                const int ow = (tc.tw * P.S + s) * kTileW + wwq;
                return ((uint32_t)(tc.img * P.Ho + oh0) * P.Wo + ow) * P.Cout + tc.nt * P.NT + cq;
            };
            auto nrows_of = [&](int s) {
                const int ow = (tc.tw * P.S + s) * kTileW + wwq;
                return ow < P.Wo ? min(4, P.Ho - oh0) : 0;        // valid tile rows of this thread (<= 0: none)
            };
            // rv[2k+i] = bias + residual of row k, channel block i (two channels)
            auto fetch = [&](int u, float2* rv) {
                if (u >= nunits) return;
                const int s = unit_s(u), grp = unit_grp(u);
                const uint32_t pix0 = pix_of(s);
                const int nrows = nrows_of(s);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int co = grp * 16 + 8 * i;
                    float2 bb = make_float2(0.f, 0.f);
                    if (P.bias && co < cw) bb = __ldg(reinterpret_cast<const float2*>(P.bias + boff + co));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float2 r = bb;
                        if (has_res && k < nrows && co < cw) {
                            const uint32_t e = pix0 + k * rowstep + co;
                            if (P.res_f) {
                                const float2 t = __ldg(reinterpret_cast<const float2*>(P.res_f + e));
                                r.x += t.x; r.y += t.y;
                            } else {
                                const float2 t = __half22float2(__ldg(reinterpret_cast<const __half2*>(P.res_hi + e)));
                                r.x += t.x; r.y += t.y;
                                if (P.res_lo) {
                                    const float2 t2 = __half22float2(__ldg(reinterpret_cast<const __half2*>(P.res_lo + e)));
                                    r.x += t2.x; r.y += t2.y;
                                }
                            }
                        }
                        rv[2 * k + i] = r;
                    }
                }
            };
            auto finish = [&](int u, const float2* rv) {
                const int s = unit_s(u), grp = unit_grp(u);
                const uint32_t pix0 = pix_of(s);
                const int nrows = nrows_of(s);
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + cs * 256 + s * P.ACC + grp * 16;
                float va[8], vb[8];
                tc_ld16x256_x2_nowait(taddr, va);
                tc_ld16x256_x2_nowait(taddr + (16u << 16), vb);
                if (P.nconcat) {
                    float wa[8], wb[8];
                    tc_ld16x256_x2_nowait(taddr + P.NT, wa);
                    tc_ld16x256_x2_nowait(taddr + (16u << 16) + P.NT, wb);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 8; ++j) { va[j] += wa[j]; vb[j] += wb[j]; }
                } else {
                    tc_wait_ld();
                }
                float2 o[2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float* v = (k < 2 ? va : vb) + 4 * i + 2 * (k & 1);
                        float2 t = make_float2(v[0] + rv[2 * k + i].x, v[1] + rv[2 * k + i].y);
                        if (P.relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); }
                        o[i][k] = t;
                    }
                if (P.y_f) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int co = grp * 16 + 8 * i;
                        if (co >= cw) continue;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < nrows) *reinterpret_cast<float2*>(P.y_f + pix0 + k * rowstep + co) = o[i][k];
                    }
                }
                if (P.y_hi) {
                    // fp16 planes: the two 8-column blocks of the group are paired up inside the quad (one shuffle with the
                    // neighbour lane per row): even lanes store 4 halves of block 0, odd lanes 4 halves of block 1 -- 8-byte
                    // stores.  Cout % 8 == 0, so a block is valid or not for the whole quad (no divergence at the shuffle).
                    const int col = grp * 16 + (odd ? 8 : 0);               // first column of the block this lane stores
                    const bool blk_ok = col < cw;
                    const uint32_t e0 = (pix0 - cq) + col + (odd ? cq - 2 : cq);
                    uint32_t ph[2][4], pl[2][4];
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t h = pack_h2_rn(o[i][k].x, o[i][k].y);
                            ph[i][k] = h;
                            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
                            pl[i][k] = pack_h2_rn(o[i][k].x - hf.x, o[i][k].y - hf.y);
                        }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? ph[0][k] : ph[1][k], 1);
                        const uint32_t lo = odd ? recv : ph[0][k], hi = odd ? ph[1][k] : recv;
                        if (blk_ok && k < nrows) *reinterpret_cast<uint2*>(P.y_hi + e0 + k * rowstep) = make_uint2(lo, hi);
                    }
                    if (P.y_lo) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? pl[0][k] : pl[1][k], 1);
                            const uint32_t lo = odd ? recv : pl[0][k], hi = odd ? pl[1][k] : recv;
                            if (blk_ok && k < nrows) *reinterpret_cast<uint2*>(P.y_lo + e0 + k * rowstep) = make_uint2(lo, hi);
                        }
                    }
                }
            };
            float2 r0[8], r1[8], r2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r0[i] = r1[i] = r2[i] = make_float2(0.f, 0.f);
            fetch(0, r0); fetch(1, r1); fetch(2, r2);
            mbar_wait(bar_acc_full + 8 * cs, (fph >> cs) & 1u);
            fph ^= 1u << cs;
            tc_fence_after();
            for (int u = 0; u < nunits; u += 3) {
                finish(u, r0); fetch(u + 3, r0);
                if (u + 1 < nunits) { finish(u + 1, r1); fetch(u + 4, r1); }
                if (u + 2 < nunits) { finish(u + 2, r2); fetch(u + 5, r2); }
            }
            tc_fence_before();
            if (P.big) { mbar_arrive(bar_acc_empty); mbar_arrive(bar_acc_empty + 8); }
            else mbar_arrive(bar_acc_empty + 8 * cs);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// weight packing: SIMT layout [wsets][ks*ks*Cin][Cout] fp32 -> swizzled smem-image blocks of split fp16.
// block (ws, nt, chunk, parity plane, tap group[, plane]) = [TG taps][rows][SWB bytes]; rows = output channels
// (nconcat: NT hi rows then NT lo rows; wsplit: a hi block followed by a lo block)
__global__ void k_pack(const Prob g, const float* __restrict__ w, __half* __restrict__ out) {
    const int blk_halves = g.b_block_bytes / 2;
    const long long total = (long long)g.wsets * g.blocks_per_set * blk_halves;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int taps = g.ks * g.ks;
    long long blk = i / blk_halves;
    const uint32_t poff = (uint32_t)(i % blk_halves) * 2;             // physical byte offset inside the block
    const uint32_t smask = g.SWB == 128 ? 7u : (g.SWB == 64 ? 3u : 1u);
    const uint32_t loff = swz(poff, smask);                            // the XOR swizzle is an involution
    const int tt = loff / g.tap_bytes;
    float v = 0.f;
    int want_lo = 0;
    if (tt < g.TG) {
        const uint32_t r = loff - tt * g.tap_bytes;
        int n = r / g.SWB; const int kk = (r % g.SWB) / 2;
        if (g.nconcat && n >= g.NT) { n -= g.NT; want_lo = 1; }
        int bi = (int)(blk % g.bpc); blk /= g.bpc;
        int slot = 0, base = 0;
        for (;;) { const int nb = g.ngrp[slot] * (g.wsplit ? 2 : 1); if (bi < base + nb || slot + 1 >= g.npa) break; base += nb; ++slot; }
        int tgi = bi - base;
        if (g.wsplit) { want_lo = tgi & 1; tgi >>= 1; }
        const int c = (int)(blk % g.nchunks); blk /= g.nchunks;
        const int nt = (int)(blk % g.ntn);
        const int ws = (int)(blk / g.ntn);
        const int k = tgi * g.TG + tt;
        if (k < g.ntap[slot]) {
            const int t = g.tapidx[slot][k];
            const int cin = c * g.KCH + kk;
            const int co = nt * g.NT + n;
            if (co < g.Cout && cin < g.Cin) v = w[((size_t)ws * taps * g.Cin + (size_t)t * g.Cin + cin) * g.Cout + co];
        }
    }
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half h = __float2half_rn(v);
    out[i] = want_lo ? __float2half_rn(v - __half2float(h)) : h;
}

// fp32 NHWC -> split-fp16 planes (test / boundary helper) and back
__global__ void k_act_split(long long n, const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = fminf(fmaxf(x[i], -65504.f), 65504.f);
    const __half h = __float2half_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2half_rn(v - __half2float(h));
}
__global__ void k_act_merge(long long n, const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = __half2float(hi[i]) + (lo ? __half2float(lo[i]) : 0.f);
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    });
    return fn;
}

// the input tensor [N][H][W][Cin] fp16 as a 4-D tensor map whose box is the largest parity-plane halo of the problem
// (all parity planes of a stride-2 problem share one map only if their boxes agree; they are encoded per slot otherwise --
//  here every slot uses the MAXIMAL box and stage_bytes is that of the maximal box, see make_prob)
static int encode_x(const Prob& g, const void* base, CUtensorMap* tm) {
    PFN_encodeTiled fn = encode_fn();
    DANET_CHECK(fn, "conv_tc: cuTensorMapEncodeTiled is not available from this driver");
    DANET_CHECK(((uintptr_t)base & 15) == 0, "conv_tc: activation plane must be 16-byte aligned");
    cuuint64_t gdim[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    cuuint64_t gstr[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.W * g.Cin * 2, (cuuint64_t)g.H * g.W * g.Cin * 2};
    const int Wb = g.sbo_a[0] / g.SWB, Hb = g.stage_bytes[0] / g.sbo_a[0];
    cuuint32_t box[4] = {(cuuint32_t)g.KCH, (cuuint32_t)(g.stride * (Wb - 1) + 1), (cuuint32_t)(g.stride * (Hb - 1) + 1), 1u};
    cuuint32_t estr[4] = {1u, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1u};
    const CUtensorMapSwizzle sw = g.SWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (g.SWB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DANET_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] box [%u,%u,%u]", (int)r,
                g.N, g.H, g.W, g.Cin, box[0], box[1], box[2]);
    return 0;
}

}  // namespace tc

static int g_sm_count[64];
static std::mutex g_tc_mu;
static unsigned long long g_tc_devs = 0;
static bool g_use_pdl = true;

int conv_tc_group_launch(int n, const danet_conv_problem* probs, cudaStream_t stream) {
    using namespace tc;
    DANET_CHECK(n >= 1 && n <= kMaxProb, "danet_conv_tc_group: 1..%d problems per launch (got %d)", kMaxProb, n);
    ArgsN a;
    memset(&a, 0, sizeof(a));
    a.nprob = n;
    int S_req[kMaxProb];
    const int s_env = env_int("DANET_TC_S", 2);
    for (int i = 0; i < n; ++i) S_req[i] = s_env;
    for (int attempt = 0;; ++attempt) {
        for (int i = 0; i < n; ++i)
            if (!make_prob(&probs[i].d, S_req[i], &a.p[i])) { set_error("danet_conv_tc_group: problem %d has an unsupported shape", i); return -1; }
        if (plan_rings(&a)) break;
        // shrink the problem with the largest halo stage to one sub-tile and retry
        int worst = -1, wb = 0;
        for (int i = 0; i < n; ++i) if (a.p[i].S > 1 && max_stage_bytes(a.p[i]) > wb) { wb = max_stage_bytes(a.p[i]); worst = i; }
        DANET_CHECK(worst >= 0 && attempt < 2 * kMaxProb, "danet_conv_tc_group: shared-memory plan does not fit");
        S_req[worst] = 1;
    }
    int base = 0;
    for (int i = 0; i < n; ++i) {
        Prob& P = a.p[i];
        const danet_conv_problem& q = probs[i];
        DANET_CHECK(q.x.hi && q.w_packed && (q.y.hi || q.y.f32), "danet_conv_tc_group: problem %d: null x.hi / weights / output", i);
        DANET_CHECK(!P.exact || q.x.lo, "danet_conv_tc_group: problem %d: exact mode needs the x.lo plane", i);
        P.wpk = (const uint8_t*)q.w_packed; P.bias = q.bias;
        P.res_f = q.res.f32; P.res_hi = (const __half*)q.res.hi; P.res_lo = (const __half*)q.res.lo;
        if (P.res_f) { P.res_hi = nullptr; P.res_lo = nullptr; }
        P.y_f = q.y.f32; P.y_hi = (__half*)q.y.hi; P.y_lo = (__half*)q.y.lo;
        if (encode_x(P, q.x.hi, &P.tm[0]) != 0) return -1;
        if (P.exact) { if (encode_x(P, q.x.lo, &P.tm[1]) != 0) return -1; }
        else P.tm[1] = P.tm[0];
        P.tile_base = base; base += P.tile_count;
        DANET_CHECK(base < (1 << 24), "danet_conv_tc_group: too many tiles");
    }
    a.total_tiles = base;
    a.prof = nullptr;
    int dev = 0;
    DANET_CUDA(cudaGetDevice(&dev));
    DANET_CHECK(dev >= 0 && dev < 64, "conv_tc: device ordinal %d out of range", dev);
    {
        std::lock_guard<std::mutex> lk(g_tc_mu);
        if (first_use_on_current_device(&g_tc_devs) != 0) {          // function attributes are per device
            DANET_CUDA(cudaDeviceGetAttribute(&g_sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
            DANET_CUDA(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            g_use_pdl = env_int("DANET_TC_PDL", 1) != 0;
        }
    }
    const int smem_bytes = kSmemFixed + a.na_stages * a.a_slot_bytes + a.nb_stages * a.b_slot_bytes;
    const int cap = g_sm_count[dev];
    const int grid = a.total_tiles < cap ? a.total_tiles : cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes < 120 * 1024 ? 120 * 1024 : smem_bytes;      // one CTA per SM (TMEM: 512 columns each)
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_use_pdl ? 1 : 0;
    DANET_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc, a));
    DANET_LAUNCH_CHECK();
    return 0;
}

}  // namespace danet

using namespace danet;

extern "C" int danet_conv_tc_supported(const danet_conv_desc* d) {
    tc::Prob g;
    if (!d || !tc::make_prob(d, 2, &g)) return 0;
    tc::ArgsN* a = new tc::ArgsN();
    a->nprob = 1; a->p[0] = g;
    bool ok = tc::plan_rings(a);
    if (!ok && tc::make_prob(d, 1, &g)) { a->p[0] = g; ok = tc::plan_rings(a); }
    delete a;
    return ok ? 1 : 0;
}

extern "C" int64_t danet_conv_tc_packed_bytes(const danet_conv_desc* d) {
    tc::Prob g;
    if (!d || !tc::make_prob(d, 1, &g)) return 0;
    return (int64_t)d->wsets * g.blocks_per_set * g.b_block_bytes;
}

extern "C" int danet_conv_tc_pack(const danet_conv_desc* d, const float* w_simt, void* w_packed, danet_stream_t stream) {
    tc::Prob g;
    DANET_CHECK(d && tc::make_prob(d, 1, &g), "danet_conv_tc_pack: shape not supported by the tcgen05 path");
    DANET_CHECK(w_simt && w_packed, "danet_conv_tc_pack: null pointer");
    const long long total = (long long)g.wsets * g.blocks_per_set * (g.b_block_bytes / 2);
    tc::k_pack<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, w_simt, (__half*)w_packed);
    DANET_LAUNCH_CHECK();
    return 0;
}

extern "C" int danet_conv_tc_group(int32_t n, const danet_conv_problem* probs, danet_stream_t stream) {
    DANET_CHECK(probs, "danet_conv_tc_group: null problem list");
    for (int i = 0; i < n; ++i) if (probs[i].d.N == 0) { DANET_CHECK(n == 1, "danet_conv_tc_group: empty problem in a group"); return 0; }
    return conv_tc_group_launch(n, probs, (cudaStream_t)stream);
}

extern "C" int danet_act_split(int64_t n, const float* x, void* hi, void* lo, danet_stream_t stream) {
    DANET_CHECK(n >= 0 && (n == 0 || (x && hi)), "danet_act_split: bad arguments");
    if (n == 0) return 0;
    tc::k_act_split<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, x, (__half*)hi, (__half*)lo);
    DANET_LAUNCH_CHECK();
    return 0;
}

extern "C" int danet_act_merge(int64_t n, const void* hi, const void* lo, float* y, danet_stream_t stream) {
    DANET_CHECK(n >= 0 && (n == 0 || (hi && y)), "danet_act_merge: bad arguments");
    if (n == 0) return 0;
    tc::k_act_merge<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, (const __half*)hi, (const __half*)lo, y);
    DANET_LAUNCH_CHECK();
    return 0;
}
