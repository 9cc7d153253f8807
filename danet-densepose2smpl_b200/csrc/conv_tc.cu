// tcgen05 implicit-GEMM convolution for sm_100a (1x1 / 3x3 / 7x7, stride 1 or 2, NHWC fp32 activations,
// fp16 tensor-core operands converted on the fly, fp32 accumulation in TMEM).
// The fast path of models/module/hr_module.py + res_module.py convolutions (conv + folded BN +
// residual + ReLU); everything it does not take goes through csrc/conv_simt.cu.
//
// Design (one persistent CTA per SM, warp-specialised, no tensor maps):
//   M tile  = 128 output pixels = 16 rows x 8 columns of one image; N tile = up to 256 output
//             channels; accumulators live in TMEM (double-buffered).
//   A (activations): the input HALO of the tile ((15*s+k) x (7*s+k) pixels) is loaded ONCE per
//             channel chunk by 8 producer warps (batched 16-byte global loads of the fp32 activations,
//             RN conversion to fp16, 16-byte st.shared) into a SWIZZLED K-major UMMA layout: one row per halo pixel,
//             SWB = 128/64/32 bytes (64/32/16 fp16 channels) per row, 16-byte chunks XOR-swizzled by the
//             row phase exactly like TMA's SWIZZLE_128B/64B/32B, halo rows padded to a pitch of
//             WP = 8k pixels so that every 8-row MMA group starts at the same swizzle phase.
//             Eight consecutive MMA rows are eight consecutive pixels of one halo row; the next
//             group is the next output row (SBO = stride * WP * SWB).  Every filter tap is only a
//             different descriptor START ADDRESS (+ base_offset = its swizzle phase), so an input
//             element crosses L2->SM ~1.4x instead of 9x (3x3).  Stride-2 convolutions store the
//             halo split by column parity so that the 8 pixels of a group stay contiguous.
//             (The first version used the no-swizzle layout: correct, but the tensor core then
//             fetches 16 bytes per cycle -- ~7 cycles per 8x16B core matrix, see profiles/.)
//   B (weights): pre-packed once (danet_conv_tc_pack) into the exact swizzled smem image of every
//             (N tile, channel chunk, tap group) block, streamed by 1-D cp.async.bulk copies that
//             complete on an mbarrier (no cuTensorMap); small weight sets stay resident in smem.
//   MMA     : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N, K=16) and
//             releases smem stages / publishes accumulators with tcgen05.commit -> mbarrier.
//   Epilogue: 8 warps read TMEM with tcgen05.ld.16x256b (a quad of lanes owns one 32-byte sector of a
//             pixel), add bias (+ residual, prefetched), ReLU, and store fp32 (or fp16 when the
//             consumer is another tensor-core convolution: danet_conv_desc.flags) NHWC.
//   Launch  : programmatic dependent launch (griddepcontrol): prologue, weight prefetch and index set-up
//             overlap the previous kernel's tail; division-free tile decode; parameter warm-up.
// Every mbarrier wait is bounded (traps instead of hanging the device).
#include "common.cuh"
#include <cuda_fp16.h>

namespace danet {
namespace tc {

constexpr int kTileH = 16, kTileW = 8;
constexpr int kThreads = 576;            // warps 0-7: A producers, warps 8-15: epilogue, warp 16: B producer, warp 17: MMA + TMEM alloc
// (the issue arbiter favours the highest warp id of a scheduler: the single MMA-issuing thread
//  must not sit behind 12 warps that poll mbarriers -- measured 370 cycles/MMA when it did)
constexpr int kWarpEpi = 8, kWarpB = 16, kWarpMma = 17;
constexpr int kPollSleepNs = 0;           // mbarrier.try_wait already suspends the warp in hardware; an extra
                                          // __nanosleep only added wake-up latency (about a microsecond per miss)
constexpr int kNumEpi = 256;             // two epilogue warps per TMEM lane quarter, each takes every other 16-column group
constexpr int kNumProducers = 256;
constexpr int kMaxBStages = 16;
constexpr int kSmemBudget = 214 * 1024;  // one CTA per SM (227 KB max - static)

struct Geom {
    int N, H, W, Cin, Cout, ks, pad, stride, relu, wsets;
    int Ho, Wo;
    int WP, HPmax;                       // A stage = one parity plane of the halo: up to HPmax rows of WP pixels (WP % 8 == 0)
    // stride-2 convolutions are decomposed by input parity (py,px): each parity plane is a dense
    // stride-1 problem with its own subset of the filter taps, so every A stage uses the full-width
    // swizzled layout.  Stride 1 = a single plane with all taps.
    int npa;                             // active parity planes (1 for stride 1, up to 4 for stride 2)
    int par_py[4], par_px[4], ntap[4], ngrp[4], blkoff[4], Hp[4], Wp[4];
    int tapoff16[4][16];                 // smem offset (16-byte units) of each tap's shifted view inside the plane
    int tapidx[4][16];                   // original filter tap index r*ks+s (for weight packing)
    int bpc;                             // weight blocks per channel chunk = sum of ngrp
    int SWB, KCH, nchunks, CGT;          // swizzle bytes per row, channels per chunk (SWB/2, fp16), ceil(Cin/KCH), 16B chunks per row
    int TG;                              // filter taps per B stage (ragged last group per plane)
    int NT, ntn;
    int tiles_w, tiles_h, total_tiles;
    int a_stage_bytes, b_stage_bytes, tap_bytes, nb_stages, na_stages;
    int smem_bytes;
    int tmem_cols, ctas_per_sm, b_resident, variant;
    int bias_smem;                       // bytes of the [wsets][Cout] bias staged in shared memory (0: read from global)
    int KS, acc_stages;                  // independent accumulators per tile (K split), TMEM accumulator stages
    int x_f16, y_f16;                    // activations in / out stored as fp16 (danet_conv_desc.flags)
    long long blocks_per_set;            // packed weight blocks per weight set
    // division-free index math: q = (x * m) >> 40 is exact for x < 2^24, divisor < 2^16 (mdiv())
    unsigned long long m_ntn, m_tw, m_th, m_ws, m_Wp[4];
    int npass[4], dhh[4], dww[4];        // A producers: passes per parity plane and the per-pass pixel step
};
static unsigned long long magic40(int d) { return (1ull << 40) / (unsigned long long)d + 1ull; }

static int tc_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DANET_TC_VARIANT"); v = e ? atoi(e) : 0; }
    return v;
}

static bool make_geom(const danet_conv_desc* d, Geom* g) {
    if (!(d->stride == 1 || d->stride == 2) || !(d->ksize == 1 || d->ksize == 3 || d->ksize == 7) || d->pad != d->ksize / 2) return false;
    if (d->Cin % 4 != 0 || d->Cout % 4 != 0 || d->H < 4 || d->W < 4) return false;
    g->N = d->N; g->H = d->H; g->W = d->W; g->Cin = d->Cin; g->Cout = d->Cout; g->ks = d->ksize;
    g->pad = d->pad; g->stride = d->stride; g->relu = d->relu; g->wsets = d->wsets;
    g->variant = tc_variant();
    g->x_f16 = (d->flags & DANET_CONV_X_F16) ? 1 : 0; g->y_f16 = (d->flags & DANET_CONV_Y_F16) ? 1 : 0;
    if ((g->x_f16 && d->Cin % 8 != 0) || (g->y_f16 && d->Cout % 8 != 0)) return false;
    g->KS = 1; g->acc_stages = 2;
    g->Ho = (d->H + 2 * d->pad - d->ksize) / d->stride + 1;
    g->Wo = (d->W + 2 * d->pad - d->ksize) / d->stride + 1;
    // parity decomposition
    g->npa = 0;
    int max_tr = 0, max_tc = 0, max_ntap = 0;
    for (int py = 0; py < d->stride; ++py)
        for (int px = 0; px < d->stride; ++px) {
            const int tr = py < d->ksize ? (d->ksize - 1 - py) / d->stride + 1 : 0;     // taps r = py + stride*i
            const int tcn = px < d->ksize ? (d->ksize - 1 - px) / d->stride + 1 : 0;
            if (tr * tcn == 0) continue;
            if (tr * tcn > 16) return false;
            const int a = g->npa++;
            g->par_py[a] = py; g->par_px[a] = px; g->ntap[a] = tr * tcn;
            g->Hp[a] = kTileH + tr - 1; g->Wp[a] = kTileW + tcn - 1;
            max_tr = tr > max_tr ? tr : max_tr; max_tc = tcn > max_tc ? tcn : max_tc;
            max_ntap = tr * tcn > max_ntap ? tr * tcn : max_ntap;
        }
    for (int a = g->npa; a < 4; ++a) { g->par_py[a] = g->par_px[a] = g->ntap[a] = g->ngrp[a] = g->blkoff[a] = g->Hp[a] = g->Wp[a] = 0; }
    g->HPmax = kTileH + max_tr - 1;
    g->WP = (kTileW + max_tc - 1 + 7) / 8 * 8;
    const int np = (d->Cout + 15) / 16 * 16;
    g->ntn = (np + 255) / 256;
    g->NT = ((np + g->ntn - 1) / g->ntn + 15) / 16 * 16;
    g->tiles_w = (g->Wo + kTileW - 1) / kTileW; g->tiles_h = (g->Ho + kTileH - 1) / kTileH;
    g->total_tiles = d->N * g->tiles_h * g->tiles_w * g->ntn;
    // Consecutive tcgen05.mma on ONE accumulator are serialised by the accumulate dependency
    // (~250-450 cycles each for these narrow N, measured); the K loop is therefore dealt round-robin
    // over KS independent TMEM accumulators that the epilogue sums.
    g->KS = 1;      // >1 deals the K loop over independent accumulators (measured: no gain; the limiter was MMA issue)
    g->acc_stages = (2 * g->KS * g->NT <= 512) ? 2 : 1;
    int cols = 32;
    while (cols < g->acc_stages * g->KS * g->NT) cols *= 2;
    g->tmem_cols = cols;
    g->bias_smem = (long long)d->wsets * d->Cout * 4 <= 8192 ? (d->wsets * d->Cout * 4 + 15) / 16 * 16 : 0;
    const int fixed = 512 + 1024 + g->bias_smem;      // barriers + 1024-byte alignment slack + staged bias
    // widest swizzle whose double-buffered halo + a minimal weight pipeline fits
    bool ok = false;
    for (int swb = 128; swb >= 32 && !ok; swb /= 2) {
        if (swb / 2 >= 2 * ((d->Cin + 15) / 16 * 16) && swb > 32) continue;   // do not pad tiny channel counts 2x to a wide row
        g->SWB = swb; g->KCH = swb / 2; g->CGT = swb / 16;      // fp16 operands: 8 channels per 16-byte chunk
        g->nchunks = (d->Cin + g->KCH - 1) / g->KCH;
        g->a_stage_bytes = g->HPmax * g->WP * swb;
        g->tap_bytes = g->NT * swb;
        for (int t = max_ntap; t >= 1 && !ok; --t) {
            if (t > 1 && t * g->tap_bytes > 32 * 1024) continue;
            g->TG = t;
            g->b_stage_bytes = (t * g->tap_bytes + 1023) / 1024 * 1024;
            if (2 * g->a_stage_bytes + 2 * g->b_stage_bytes + fixed <= kSmemBudget) ok = true;
        }
    }
    if (!ok) return false;
    g->bpc = 0;
    for (int a = 0; a < g->npa; ++a) {
        g->ngrp[a] = (g->ntap[a] + g->TG - 1) / g->TG;
        g->blkoff[a] = g->bpc; g->bpc += g->ngrp[a];
        const int tcn = g->Wp[a] - kTileW + 1;
        for (int k = 0; k < 16; ++k) { g->tapoff16[a][k] = 0; g->tapidx[a][k] = 0; }
        for (int k = 0; k < g->ntap[a]; ++k) {
            const int ti = k / tcn, tj = k % tcn;
            g->tapoff16[a][k] = ((ti * g->WP + tj) * g->SWB) >> 4;
            g->tapidx[a][k] = (g->par_py[a] + d->stride * ti) * d->ksize + (g->par_px[a] + d->stride * tj);
        }
    }
    const int nblk = g->nchunks * g->bpc;
    g->na_stages = 2;
    g->b_resident = 0; g->ctas_per_sm = 1;
    if (d->wsets == 1 && g->ntn == 1 && nblk <= kMaxBStages &&
        fixed + 2 * g->a_stage_bytes + nblk * g->b_stage_bytes <= kSmemBudget && g->total_tiles >= 2 * 148) {
        g->b_resident = 1; g->nb_stages = nblk;           // the whole weight set stays in shared memory
        if (fixed + 3 * g->a_stage_bytes + nblk * g->b_stage_bytes <= kSmemBudget) g->na_stages = 3;
    } else {
        int nb = (kSmemBudget - fixed - 2 * g->a_stage_bytes) / g->b_stage_bytes;
        g->nb_stages = nb > kMaxBStages ? kMaxBStages : nb;
        if (g->nb_stages >= 6 && fixed + 3 * g->a_stage_bytes + 4 * g->b_stage_bytes <= kSmemBudget) {
            g->na_stages = 3;
            nb = (kSmemBudget - fixed - 3 * g->a_stage_bytes) / g->b_stage_bytes;
            g->nb_stages = nb > kMaxBStages ? kMaxBStages : nb;
        }
    }
    g->smem_bytes = fixed + g->na_stages * g->a_stage_bytes + g->nb_stages * g->b_stage_bytes;
    if (g->smem_bytes < 120 * 1024) g->smem_bytes = 120 * 1024;       // one CTA per SM (TMEM budget)
    g->blocks_per_set = (long long)g->ntn * nblk;
    if (g->total_tiles >= (1 << 24) || g->wsets >= (1 << 16)) return false;
    if ((long long)d->N * g->Ho * g->Wo * d->Cout >= (1LL << 31) || (long long)d->H * d->W * d->Cin >= (1LL << 31)) return false;   // 32-bit element offsets
    g->m_ntn = magic40(g->ntn); g->m_tw = magic40(g->tiles_w); g->m_th = magic40(g->tiles_h); g->m_ws = magic40(g->wsets);
    {
        const int ppt = kNumProducers / g->CGT;
        for (int a = 0; a < 4; ++a) {
            const int Wp = g->Wp[a] > 0 ? g->Wp[a] : 1;
            g->m_Wp[a] = magic40(Wp);
            g->npass[a] = (g->Hp[a] * g->Wp[a] + ppt - 1) / ppt;
            g->dhh[a] = ppt / Wp; g->dww[a] = ppt - g->dhh[a] * Wp;
        }
    }
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
// same, but backs off between polls so that waiting warps do not steal issue slots from the MMA thread
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (kPollSleepNs > 0) __nanosleep(kPollSleepNs);
    }
    __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// programmatic dependent launch: this grid may start while the previous kernel of the stream drains;
// nothing the previous kernel wrote (activations, residual) or still reads (our output buffer may be
// its input) is touched before pdl_wait()
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // d = {hi -> upper, lo -> lower}
    return r;
}
// K-major swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1).
// layout_type: SWIZZLE_128B = 2, SWIZZLE_64B = 4, SWIZZLE_32B = 6; LBO is unused for swizzled K-major.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type, uint32_t base_offset) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((base_offset & 7u) << 17) | (layout_type << 29);
    return ((uint64_t)hi << 32) | lo;
}
// physical offset of logical byte offset `off` inside a 1024-byte-aligned swizzled tile
__device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t mask) { return off ^ (((off >> 7) & mask) << 4); }

__device__ __forceinline__ int mdiv(int x, unsigned long long m) { return (int)(((unsigned long long)(unsigned)x * m) >> 40); }
// tile -> (N tile, tile column, tile row, image)
__device__ __forceinline__ void decode_tile(const Geom& g, int tile, int& nt, int& tw, int& th, int& img) {
    int r = mdiv(tile, g.m_ntn); nt = tile - r * g.ntn;
    int r2 = mdiv(r, g.m_tw); tw = r - r2 * g.tiles_w;
    img = mdiv(r2, g.m_th); th = r2 - img * g.tiles_h;
}

#define TC_PROF_BEGIN() long long _t0 = prof_on ? clock64() : 0
#define TC_PROF_END(slot) do { if (prof_on) prof_acc[slot] += clock64() - _t0; } while (0)

struct Args {
    Geom g;
    const void* x; const float* wpk; const float* bias; const float* res; void* y;   // x / y: fp32, or fp16 when g.x_f16 / g.y_f16
    long long* prof;      // optional [16] cycle counters of CTA 0 (bring-up instrumentation), else NULL
};

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// PROF: bring-up instrumentation; RES: a residual tensor is added in the epilogue (compile-time so that the
// layers without one carry none of the prefetch code)
template <bool PROF, bool RES>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_tc(const Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Geom& g = a.g;
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;          // swizzle atoms need 1024-byte alignment
    const uint32_t sA = sbase;
    const uint32_t sB = sbase + g.na_stages * g.a_stage_bytes;
    const uint32_t sBar = sB + g.nb_stages * g.b_stage_bytes;
    // barrier map (8 bytes each)
    const uint32_t bar_a_full = sBar, bar_a_empty = sBar + 24, bar_acc_full = sBar + 48, bar_acc_empty = sBar + 64;
    const uint32_t bar_b_full = sBar + 80, bar_b_empty = sBar + 80 + 8 * kMaxBStages;
    const uint32_t tmem_slot_addr = sBar + 80 + 16 * kMaxBStages;       // 80 + 256 + 4 <= 512
    const uint32_t sBias = sBar + 512;                                  // [wsets][Cout] fp32 bias (when g.bias_smem)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the kernel parameters (about 1 KB, new for every launch) are read through the constant cache:
    // touch every word once, all misses in flight together, before the roles need them one by one
    // (measured 1.4 us of dependent cold misses + divisions before the first global load, profiles/)
    if (threadIdx.x < (int)(sizeof(Args) / 4)) {
        const int w = reinterpret_cast<const int*>(&a)[threadIdx.x];
        if (w == 0x7fffdead) reinterpret_cast<volatile int*>(smem)[threadIdx.x] = w;
    }
    const long long t_entry = clock64();
    long long* tl = (PROF && a.prof) ? a.prof + 16 + 16 * blockIdx.x : nullptr;       // per-CTA timeline (bring-up)
    if (PROF && tl && threadIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); tl[0] = (long long)gt; }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar_acc_full + 8 * i, 1); mbar_init(bar_acc_empty + 8 * i, kNumEpi); }   // stage 1 unused when acc_stages == 1
        for (int i = 0; i < g.na_stages; ++i) { mbar_init(bar_a_full + 8 * i, kNumProducers); mbar_init(bar_a_empty + 8 * i, 1); }
        for (int i = 0; i < g.nb_stages; ++i) { mbar_init(bar_b_full + 8 * i, 1); mbar_init(bar_b_empty + 8 * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (g.bias_smem && a.bias)                                          // weights: independent of the previous kernel
        for (int i = threadIdx.x; i < g.wsets * g.Cout; i += kThreads)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sBias + 4 * i), "f"(__ldg(a.bias + i)) : "memory");
    if (warp == kWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot_addr), "r"((uint32_t)g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot_addr));

    const int HWC = g.H * g.W;
    pdl_launch_dependents();            // the next launch may fill SMs as our CTAs retire
    if (PROF && tl && threadIdx.x == 0) tl[2] = clock64() - t_entry;
    if (warp == kWarpB) {
        // ================= B producer: bulk copies of pre-packed weight blocks =================
        if (lane == 0) {
            int bs = 0; uint32_t bph = 0;
            for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
                int nt, tw_, th_, img;
                decode_tile(g, tile, nt, tw_, th_, img);
                const int ws = img - mdiv(img, g.m_ws) * g.wsets;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wpk) +
                    ((long long)ws * g.blocks_per_set + (long long)nt * g.nchunks * g.bpc) * g.b_stage_bytes;
                const int nblk = g.nchunks * g.bpc;
                if (g.b_resident && tile != (int)blockIdx.x) break;          // weights already resident
                for (int b = 0; b < nblk; ++b) {
                    if (!g.b_resident) mbar_wait_sleep(bar_b_empty + 8 * bs, bph ^ 1);
                    if (PROF && tl && b == 0 && tile == (int)blockIdx.x) tl[12] = clock64() - t_entry;
                    mbar_expect_tx(bar_b_full + 8 * bs, g.b_stage_bytes);
                    bulk_g2s(sB + bs * g.b_stage_bytes, src + (long long)b * g.b_stage_bytes, g.b_stage_bytes, bar_b_full + 8 * bs);
                    if (++bs == g.nb_stages) { bs = 0; bph ^= 1; }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ================= MMA issuer =================
        // The whole warp runs the loop (warp-uniform control flow keeps descriptors in uniform
        // registers); one elected lane issues the tcgen05 instructions.  A divergent `lane == 0`
        // region made ptxas wrap every UTCHMMA in an ELECT/BRA.U.ANY loop: ~30 SASS instructions
        // and ~240 cycles per MMA (profiles/r01_ncu_conv_tc_v6_summary.txt).
        {
            // kind::f16, A/B = F16 (format 0), D = F32, both K-major, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | ((uint32_t)(g.NT >> 3) << 17) | ((128u >> 4) << 24);
            int as = 0, bs = 0, cs = 0; uint32_t aph = 0, bph = 0, cph = 0;
            const uint32_t sbo_a = g.WP * g.SWB, sbo_b = 8 * g.SWB;
            const uint32_t ltype = g.SWB == 128 ? 2u : (g.SWB == 64 ? 4u : 6u);
            // 64-bit descriptors are advanced by plain adds on their address field (16-byte units):
            // all shared-memory addresses are < 256 KB, so the 14-bit field never carries.
            const uint64_t ad0 = make_desc(0, sbo_a, ltype, 0u);
            const uint64_t bd0 = make_desc(0, sbo_b, ltype, 0u);
            const int kmma = g.KCH / 16;                         // K = 16 halves (32 bytes) per MMA
            const uint32_t tap16 = g.tap_bytes >> 4;
            bool first_tile = true;
            const bool prof_on = PROF && a.prof != nullptr && blockIdx.x == 0;
            long long prof_acc[4] = {0, 0, 0, 0};
            const long long t_start = prof_on ? clock64() : 0;
            for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
                { TC_PROF_BEGIN(); mbar_wait(bar_acc_empty + 8 * cs, cph ^ 1); TC_PROF_END(0); }
                tc_fence_after();
                const uint32_t d_base = tmem_base + cs * g.NT;
                uint32_t acc = 0;
                for (int c = 0, u = 0; c < g.nchunks; ++c)
                for (int slot = 0; slot < g.npa; ++slot, ++u) {
                    { TC_PROF_BEGIN(); mbar_wait(bar_a_full + 8 * as, aph); TC_PROF_END(1); }
                    if (PROF && tl && first_tile && u == 0 && lane == 0) tl[4] = clock64() - t_entry;
                    fence_proxy_async();
                    tc_fence_after();
                    const uint64_t ad_st = ad0 + ((sA + as * g.a_stage_bytes) >> 4);
                    const int kreal = (g.Cin - c * g.KCH + 15) >> 4;
                    const int kv = kreal < kmma ? kreal : kmma;
                    for (int tg = 0; tg < g.ngrp[slot]; ++tg) {
                        if (!g.b_resident || first_tile) {
                            TC_PROF_BEGIN();
                            mbar_wait(bar_b_full + 8 * bs, g.b_resident ? 0u : bph);
                            tc_fence_after();
                            TC_PROF_END(2);
                            if (PROF && tl && first_tile && u == 0 && tg == 0 && lane == 0) tl[3] = clock64() - t_entry;
                        }
                        uint64_t bd = bd0 + ((sB + bs * g.b_stage_bytes) >> 4);
                        const int k0 = tg * g.TG;
                        const int ntk = min(g.TG, g.ntap[slot] - k0);
                        if (elect_one()) {
                            // K steps whose 16 channels lie entirely beyond Cin hold zeros in A and B: not issued
                            if (kv == 4) {
                                for (int tt = 0; tt < ntk; ++tt) {
                                    const uint64_t ad = ad_st + (uint32_t)g.tapoff16[slot][k0 + tt];
                                    tc_mma_f16(d_base, ad, bd, idesc, acc);
                                    tc_mma_f16(d_base, ad + 2, bd + 2, idesc, 1u);
                                    tc_mma_f16(d_base, ad + 4, bd + 4, idesc, 1u);
                                    tc_mma_f16(d_base, ad + 6, bd + 6, idesc, 1u);
                                    acc = 1; bd += tap16;
                                }
                            } else if (kv == 3) {
                                for (int tt = 0; tt < ntk; ++tt) {
                                    const uint64_t ad = ad_st + (uint32_t)g.tapoff16[slot][k0 + tt];
                                    tc_mma_f16(d_base, ad, bd, idesc, acc);
                                    tc_mma_f16(d_base, ad + 2, bd + 2, idesc, 1u);
                                    tc_mma_f16(d_base, ad + 4, bd + 4, idesc, 1u);
                                    acc = 1; bd += tap16;
                                }
                            } else if (kv == 2) {
                                for (int tt = 0; tt < ntk; ++tt) {
                                    const uint64_t ad = ad_st + (uint32_t)g.tapoff16[slot][k0 + tt];
                                    tc_mma_f16(d_base, ad, bd, idesc, acc);
                                    tc_mma_f16(d_base, ad + 2, bd + 2, idesc, 1u);
                                    acc = 1; bd += tap16;
                                }
                            } else {
                                for (int tt = 0; tt < ntk; ++tt) {
                                    tc_mma_f16(d_base, ad_st + (uint32_t)g.tapoff16[slot][k0 + tt], bd, idesc, acc);
                                    acc = 1; bd += tap16;
                                }
                            }
                            if (!g.b_resident) tc_commit(bar_b_empty + 8 * bs);
                        }
                        __syncwarp();
                        acc = 1;
                        if (++bs == g.nb_stages) { bs = 0; bph ^= 1; }
                    }
                    if (elect_one()) tc_commit(bar_a_empty + 8 * as);
                    __syncwarp();
                    if (++as == g.na_stages) { as = 0; aph ^= 1; }
                }
                if (elect_one()) tc_commit(bar_acc_full + 8 * cs);
                __syncwarp();
                if (PROF && tl && lane == 0) tl[5] = clock64() - t_entry;
                if (++cs == g.acc_stages) { cs = 0; cph ^= 1; }
                first_tile = false;
            }
            if (prof_on && lane == 0) {
                a.prof[0] = clock64() - t_start; a.prof[1] = prof_acc[0]; a.prof[2] = prof_acc[1]; a.prof[3] = prof_acc[2];
            }
        }
    } else if (warp < kWarpEpi) {
        // ================= A producers: halo tile -> smem (swizzled K-major rows) =================
        // thread <-> (channel group cg, pixel slot); pixels advance by a fixed step per pass so the
        // halo coordinates are updated incrementally (no divisions in the loop); every pass's
        // global load is issued before the first shared store (one latency exposure per 8 passes).
        const int pt = threadIdx.x;                             // 0..255
        const int cg = pt & (g.CGT - 1);                         // 16-byte chunk (8 fp16 channels) within the row
        const int px0 = pt / g.CGT;
        const uint32_t smask = g.SWB == 128 ? 7u : (g.SWB == 64 ? 3u : 1u);
        int as = 0; uint32_t aph = 0;
        pdl_wait();                                              // activations come from the previous kernel
        const bool prof_on = PROF && a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
        long long prof_acc[1] = {0};
        const long long t_start = prof_on ? clock64() : 0;
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
            int nt_, tw, th, img;
            decode_tile(g, tile, nt_, tw, th, img);
            const int h0 = th * kTileH * g.stride - g.pad, w0 = tw * kTileW * g.stride - g.pad;
            const float* xi = reinterpret_cast<const float*>(a.x) + (size_t)img * HWC * g.Cin + cg * 8;
            const __half* xi16 = reinterpret_cast<const __half*>(a.x) + (size_t)img * HWC * g.Cin + cg * 8;
            for (int c = 0, u = 0; c < g.nchunks; ++c)
            for (int slot = 0; slot < g.npa; ++slot, ++u) {
                const int Wp = g.Wp[slot], Hp = g.Hp[slot];
                const int hb = h0 + g.par_py[slot], wb = w0 + g.par_px[slot];
                const int npass = g.npass[slot];
                const int dhh = g.dhh[slot], dww = g.dww[slot];
                int hh = mdiv(px0, g.m_Wp[slot]), ww = px0 - hh * Wp;
                if (PROF && tl && pt == 0 && tile == (int)blockIdx.x && u == 0) tl[8] = clock64() - t_entry;
                { TC_PROF_BEGIN(); mbar_wait_sleep(bar_a_empty + 8 * as, aph ^ 1); TC_PROF_END(0); }
                const uint32_t a_st = sA + as * g.a_stage_bytes;
                const bool ch_ok = c * g.KCH + cg * 8 < g.Cin;       // channels beyond Cin are zero-filled in smem
                const bool ch_ok2 = c * g.KCH + cg * 8 + 4 < g.Cin;  // second float4 of the 8-channel chunk
                const float* xc = xi + c * g.KCH;
                if (g.x_f16) {
                    // fp16 activations: every 16-byte piece (8 channels) is one cp.async with zero fill, the
                    // stage barrier is armed by cp.async.mbarrier.arrive.noinc: no registers, no conversion
                    const __half* xc16 = xi16 + c * g.KCH;
                    for (int p = 0; p < npass; ++p) {
                        if (hh < Hp) {
                            const int ih = hb + g.stride * hh, iw = wb + g.stride * ww;
                            const uint32_t dst = a_st + swz((uint32_t)(hh * g.WP + ww) * g.SWB + cg * 16, smask);
                            const bool ok = ch_ok && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
                            const void* src = ok ? (const void*)(xc16 + (uint32_t)((ih * g.W + iw) * g.Cin)) : a.x;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
                        }
                        ww += dww; hh += dhh;
                        if (ww >= Wp) { ww -= Wp; hh += 1; }
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_a_full + 8 * as) : "memory");
                    if (++as == g.na_stages) { as = 0; aph ^= 1; }
                    continue;
                }
                // fp32 activations are converted to fp16 (RN, saturating) on the way into shared memory:
                // half the operand bytes per MAC for the tensor core and twice the K per MMA.  All global
                // loads of a batch are issued before the first conversion/store.
                for (int p0 = 0; p0 < npass; p0 += 6) {
                    float4 v0[6], v1[6];
                    uint32_t dst[6];
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        v0[q] = make_float4(0.f, 0.f, 0.f, 0.f); v1[q] = v0[q];
                        dst[q] = 0xFFFFFFFFu;
                        if (p0 + q < npass && hh < Hp) {
                            const int ih = hb + g.stride * hh, iw = wb + g.stride * ww;
                            dst[q] = a_st + swz((uint32_t)(hh * g.WP + ww) * g.SWB + cg * 16, smask);
                            if (ch_ok && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) {
                                const float4* src = reinterpret_cast<const float4*>(xc + (uint32_t)((ih * g.W + iw) * g.Cin));
                                v0[q] = __ldg(src);
                                if (ch_ok2) v1[q] = __ldg(src + 1);
                            }
                        }
                        ww += dww; hh += dhh;
                        if (ww >= Wp) { ww -= Wp; hh += 1; }
                    }
                    if (PROF && tl && pt == 0 && tile == (int)blockIdx.x && u == 0 && p0 == 0) { tl[9] = clock64() - t_entry; if (__float_as_uint(v0[0].x) == 0x12345u) tl[15] = 1; tl[10] = clock64() - t_entry; }
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        if (dst[q] != 0xFFFFFFFFu)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst[q]), "r"(pack_h2(v0[q].x, v0[q].y)),
                                         "r"(pack_h2(v0[q].z, v0[q].w)), "r"(pack_h2(v1[q].x, v1[q].y)), "r"(pack_h2(v1[q].z, v1[q].w)) : "memory");
                    }
                }
                fence_proxy_async();
                mbar_arrive(bar_a_full + 8 * as);
                if (PROF && tl && pt == 0 && tile == (int)blockIdx.x && u == 0) tl[11] = clock64() - t_entry;
                if (++as == g.na_stages) { as = 0; aph ^= 1; }
            }
        }
        if (prof_on) { a.prof[4] = clock64() - t_start; a.prof[5] = prof_acc[0]; }
    } else if (warp < kWarpB) {
        // ================= epilogue: TMEM -> bias/residual/ReLU -> global =================
        // Each of the two warps of a TMEM lane quarter takes every other 16-column group.  The residual
        // operands of up to THREE groups ahead are held in registers and the first three are requested
        // BEFORE the accumulator wait: one group iteration used to cost a full global-load latency
        // (~1.5 us, profiles/r01_tc_role_cycles_v11.log).  (Two smem-staged, fully coalesced variants --
        // block-wide slabs and warp-private slabs with two alternating warp groups -- were both measured
        // SLOWER: profiles/r01_tc_role_cycles_v12_*, _v14_*; the lane = row mapping of the first versions
        // cost one L1 line per 16 bytes.)
        const int q = warp & 3;                                  // TMEM lane quarter this warp may access
        const int half = (warp - kWarpEpi) >> 2;                 // 0/1: which 16-column groups of the tile this warp owns
        const int ngroups = g.NT / 16;
        int cs = 0; uint32_t cph = 0;
        pdl_wait();                                              // residual reads / output writes
        const bool prof_on = PROF && a.prof != nullptr && blockIdx.x == 0 && warp == kWarpEpi && lane == 0;
        long long prof_acc[1] = {0};
        long long ph[6] = {0, 0, 0, 0, 0, 0};      // PROF: decode+addresses, residual fetch issue, acc wait, tcgen05.ld, finish (math+stores), arrive
        const long long t_start = prof_on ? clock64() : 0;
        {
            // Quad mapping (tcgen05.ld 16x256b): a thread owns tile column ww = lane/4 of the four tile rows
            // 4q..4q+3 and, per 16-column group, channels 2*(lane%4)+{0,1} and +8: the four lanes of a quad
            // read/write one whole 32-byte sector, a warp instruction touches 8 lines instead of 32.  The
            // lane = row mapping (16 bytes per line per instruction) made the load/store pipe the limiter of
            // the epilogue: 8-9 us for one exposed 128 x 192 tile (per-CTA timeline in profiles/).
            const int wwq = lane >> 2, cq = 2 * (lane & 3);
            for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
                long long tp = prof_on ? clock64() : 0;
                int nt, tw, th, img;
                decode_tile(g, tile, nt, tw, th, img);
                const int ow = tw * kTileW + wwq, oh0 = th * kTileH + 4 * q;
                // element offsets fit 32 bits (make_geom refuses tensors of 2^31 elements or more)
                const uint32_t pix0 = ((uint32_t)(img * g.Ho + oh0) * g.Wo + ow) * g.Cout + nt * g.NT + cq;
                const uint32_t rowstep = (uint32_t)(g.Wo * g.Cout);
                const int nrows = ow < g.Wo ? min(4, g.Ho - oh0) : 0;        // valid tile rows of this thread (<= 0: none)
                const int chlim = g.Cout - nt * g.NT - cq;                   // channel offsets below this are real
                const int boff = (img - mdiv(img, g.m_ws) * g.wsets) * g.Cout + nt * g.NT + cq;   // first channel of this thread
                const float* bias = a.bias ? a.bias + boff : nullptr;
                constexpr bool has_res = RES;
                // (without a residual the whole prefetch is skipped: its predicate arithmetic alone cost ~500 cycles
                //  per tile, profiles/r01_tc_epilogue_phases.txt)
                auto fetch = [&](int grp, float2* rv) {
                    if (!has_res) return;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            rv[2 * k + i] = make_float2(0.f, 0.f);
                            const int co = grp * 16 + 8 * i;
                            if (grp < ngroups && k < nrows && co < chlim)
                                rv[2 * k + i] = __ldg(reinterpret_cast<const float2*>(a.res + pix0 + k * rowstep + co));
                        }
                };
                auto finish = [&](int grp, const float* va, const float* vb, const float2* rv) {
                    if (g.y_f16) {
                        // fp16 output: the two 8-column blocks of the group are paired up inside the quad (one
                        // shuffle with the neighbour lane per row): even lanes store 4 halves of block 0, odd lanes
                        // 4 halves of block 1 -- 8-byte stores, half the store instructions and L1 lines of 4-byte ones.
                        // Cout % 8 == 0 here, so a block is valid or not for the whole quad (no divergence at the shuffle).
                        const bool odd = lane & 1;
                        const int cw = g.Cout - nt * g.NT;                       // channels of this N tile that exist
                        uint32_t pk[2][4];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int co = grp * 16 + 8 * i;
                            float2 bb = make_float2(0.f, 0.f);
                            if (bias && co < cw) {
                                if (g.bias_smem) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(bb.x), "=f"(bb.y) : "r"(sBias + 4 * (boff + co)));
                                else bb = __ldg(reinterpret_cast<const float2*>(bias + co));
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float* v = (k < 2 ? va : vb) + 4 * i + 2 * (k & 1);
                                float2 o = make_float2(v[0] + bb.x + rv[2 * k + i].x, v[1] + bb.y + rv[2 * k + i].y);
                                if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); }
                                pk[i][k] = pack_h2(o.x, o.y);
                            }
                        }
                        const int col = grp * 16 + (odd ? 8 : 0);               // first column of the block this lane stores
                        const bool blk_ok = col < cw;
                        __half* yrow = reinterpret_cast<__half*>(a.y) + (pix0 - cq) + col + (odd ? cq - 2 : cq);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? pk[0][k] : pk[1][k], 1);
                            const uint32_t lo = odd ? recv : pk[0][k], hi = odd ? pk[1][k] : recv;
                            if (blk_ok && k < nrows) *reinterpret_cast<uint2*>(yrow + k * rowstep) = make_uint2(lo, hi);
                        }
                        return;
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int co = grp * 16 + 8 * i;
                        if (co >= chlim) continue;
                        float2 bb = make_float2(0.f, 0.f);
                        // a global bias load sat in the dependent chain of every group: an L2 round trip each
                        // time its line had been evicted (4.3K cycles per 128 x 64 tile in the 1x1 24->64 layer)
                        if (bias) {
                            if (g.bias_smem) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(bb.x), "=f"(bb.y) : "r"(sBias + 4 * (boff + co)));
                            else bb = __ldg(reinterpret_cast<const float2*>(bias + co));
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k >= nrows) continue;
                            const float* v = (k < 2 ? va : vb) + 4 * i + 2 * (k & 1);
                            float2 o = make_float2(v[0] + bb.x + rv[2 * k + i].x, v[1] + bb.y + rv[2 * k + i].y);
                            if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); }
                            *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.y) + pix0 + k * rowstep + co) = o;
                        }
                    }
                };
                float2 r0[8], r1[8], r2[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) r0[i] = r1[i] = r2[i] = make_float2(0.f, 0.f);
                if (prof_on) { const long long t = clock64(); ph[0] += t - tp; tp = t; }
                fetch(half, r0); fetch(half + 2, r1); fetch(half + 4, r2);
                if (prof_on) { const long long t = clock64(); ph[1] += t - tp; tp = t; }
                { TC_PROF_BEGIN(); mbar_wait_sleep(bar_acc_full + 8 * cs, cph); TC_PROF_END(0); }
                if (prof_on) { const long long t = clock64(); ph[2] += t - tp; tp = t; }
                if (PROF && tl && warp == kWarpEpi && lane == 0 && tile == (int)blockIdx.x) tl[6] = clock64() - t_entry;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + cs * g.NT;
                if constexpr (!RES) {
                    // no residual registers to carry: TMEM loads of TWO column groups are in flight before the one
                    // tcgen05.wait::ld (the wait covers every outstanding load anyway)
                    for (int grp = half; grp < ngroups; grp += 4) {
                        float va[8], vb[8], wa[8], wb[8];
                        const bool two = grp + 2 < ngroups;
                        tc_ld16x256_x2_nowait(taddr + grp * 16, va);
                        tc_ld16x256_x2_nowait(taddr + (16u << 16) + grp * 16, vb);
                        if (two) {
                            tc_ld16x256_x2_nowait(taddr + (grp + 2) * 16, wa);
                            tc_ld16x256_x2_nowait(taddr + (16u << 16) + (grp + 2) * 16, wb);
                        }
                        tc_wait_ld();
                        if (prof_on) { const long long t = clock64(); ph[3] += t - tp; tp = t; }
                        finish(grp, va, vb, r0);
                        if (two) finish(grp + 2, wa, wb, r0);
                        if (prof_on) { const long long t = clock64(); ph[4] += t - tp; tp = t; }
                    }
                } else
                for (int grp = half; grp < ngroups; grp += 6) {
                    float va[8], vb[8];
                    tc_ld16x256_x2_nowait(taddr + grp * 16, va);
                    tc_ld16x256_x2_nowait(taddr + (16u << 16) + grp * 16, vb);
                    tc_wait_ld();
                    if (prof_on) { const long long t = clock64(); ph[3] += t - tp; tp = t; }
                    finish(grp, va, vb, r0);
                    fetch(grp + 6, r0);
                    if (prof_on) { const long long t = clock64(); ph[4] += t - tp; tp = t; }
                    if (grp + 2 < ngroups) {
                        tc_ld16x256_x2_nowait(taddr + (grp + 2) * 16, va);
                        tc_ld16x256_x2_nowait(taddr + (16u << 16) + (grp + 2) * 16, vb);
                        tc_wait_ld();
                        if (prof_on) { const long long t = clock64(); ph[3] += t - tp; tp = t; }
                        finish(grp + 2, va, vb, r1);
                        fetch(grp + 8, r1);
                        if (prof_on) { const long long t = clock64(); ph[4] += t - tp; tp = t; }
                    }
                    if (grp + 4 < ngroups) {
                        tc_ld16x256_x2_nowait(taddr + (grp + 4) * 16, va);
                        tc_ld16x256_x2_nowait(taddr + (16u << 16) + (grp + 4) * 16, vb);
                        tc_wait_ld();
                        finish(grp + 4, va, vb, r2);
                        fetch(grp + 10, r2);
                    }
                }
                tc_fence_before();
                mbar_arrive(bar_acc_empty + 8 * cs);
                if (++cs == g.acc_stages) { cs = 0; cph ^= 1; }
                if (prof_on) { const long long t = clock64(); ph[5] += t - tp; tp = t; }
            }
        }
        if (prof_on) for (int i = 0; i < 6; ++i) a.prof[8 + i] = ph[i];
        if (prof_on) { a.prof[6] = clock64() - t_start; a.prof[7] = prof_acc[0]; }
        if (PROF && tl && warp == kWarpEpi && lane == 0) tl[7] = clock64() - t_entry;
    }
    tc_fence_before();
    __syncthreads();
    if (PROF && tl && threadIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); tl[1] = (long long)gt; }
    if (warp == kWarpMma) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    }
}

// weight packing: SIMT layout [wsets][ks*ks*Cin][Cout] fp32 -> swizzled smem-image blocks of fp16
// block (ws, nt, chunk, tap group) = [TG taps][NT rows][SWB bytes], rows = output channels
__global__ void k_pack(const Geom g, const float* __restrict__ w, __half* __restrict__ out) {
    const int blk_halves = g.b_stage_bytes / 2;
    const long long total = (long long)g.wsets * g.blocks_per_set * blk_halves;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int taps = g.ks * g.ks;
    long long blk = i / blk_halves;
    const uint32_t poff = (uint32_t)(i % blk_halves) * 2;             // physical byte offset inside the block
    const int tt = poff / g.tap_bytes;
    float v = 0.f;
    if (tt < g.TG) {
        const uint32_t smask = g.SWB == 128 ? 7u : (g.SWB == 64 ? 3u : 1u);
        const uint32_t loff = swz(poff - tt * g.tap_bytes, smask);     // the XOR swizzle is an involution
        const int n = loff / g.SWB, kk = (loff % g.SWB) / 2;
        int bi = (int)(blk % g.bpc); blk /= g.bpc;
        int slot = 0;
        while (slot + 1 < g.npa && bi >= g.blkoff[slot + 1]) ++slot;
        const int tgi = bi - g.blkoff[slot];
        const int c = (int)(blk % g.nchunks); blk /= g.nchunks;
        const int nt = (int)(blk % g.ntn);
        const int ws = (int)(blk / g.ntn);
        const int k = tgi * g.TG + tt;
        if (k >= g.ntap[slot]) { out[i] = __float2half_rn(0.f); return; }
        const int t = g.tapidx[slot][k];
        const int cin = c * g.KCH + kk;
        const int co = nt * g.NT + n;
        if (co < g.Cout && cin < g.Cin) v = w[((size_t)ws * taps * g.Cin + (size_t)t * g.Cin + cin) * g.Cout + co];
    }
    out[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
}

}  // namespace tc

static long long* g_tc_prof = nullptr;

int conv_tc_launch(const danet_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                   const float* residual, void* y, cudaStream_t stream) {
    tc::Args a;
    if (!tc::make_geom(d, &a.g)) { set_error("conv_tc_launch: unsupported shape"); return -1; }
    a.x = x; a.wpk = (const float*)w_packed; a.bias = bias; a.res = residual; a.y = y;
    a.prof = g_tc_prof;
    static int sm_count_of[64];
    static unsigned long long attr_devs = 0;
    static bool use_pdl = true;
    int dev = 0;
    DANET_CUDA(cudaGetDevice(&dev));
    DANET_CHECK(dev >= 0 && dev < 64, "conv_tc_launch: device ordinal %d out of range", dev);
    if (first_use_on_current_device(&attr_devs) != 0) {          // function attributes are per device
        DANET_CUDA(cudaDeviceGetAttribute(&sm_count_of[dev], cudaDevAttrMultiProcessorCount, dev));
        DANET_CUDA(cudaFuncSetAttribute(tc::k_conv_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        DANET_CUDA(cudaFuncSetAttribute(tc::k_conv_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        DANET_CUDA(cudaFuncSetAttribute(tc::k_conv_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        DANET_CUDA(cudaFuncSetAttribute(tc::k_conv_tc<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        const char* e = getenv("DANET_TC_PDL");
        use_pdl = !(e && atoi(e) == 0);
    }
    const int cap = sm_count_of[dev] * a.g.ctas_per_sm;
    a.g.variant = tc::tc_variant();
    const int grid = a.g.total_tiles < cap ? a.g.total_tiles : cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tc::kThreads); cfg.dynamicSmemBytes = a.g.smem_bytes; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
    if (a.prof) {
        if (a.res) { DANET_CUDA(cudaLaunchKernelEx(&cfg, tc::k_conv_tc<true, true>, a)); }
        else { DANET_CUDA(cudaLaunchKernelEx(&cfg, tc::k_conv_tc<true, false>, a)); }
    } else {
        if (a.res) { DANET_CUDA(cudaLaunchKernelEx(&cfg, tc::k_conv_tc<false, true>, a)); }
        else { DANET_CUDA(cudaLaunchKernelEx(&cfg, tc::k_conv_tc<false, false>, a)); }
    }
    DANET_LAUNCH_CHECK();
    return 0;
}

}  // namespace danet

using namespace danet;

// bring-up instrumentation: device buffer of 16 int64 cycle counters written by CTA 0 of every
// subsequent tcgen05 conv launch ([0] MMA warp total, [1] wait acc_empty, [2] wait A_full,
// [3] wait B_full, [4] producer total, [5] producer wait A_empty, [6] epilogue total,
// [7] epilogue wait acc_full); NULL disables it
extern "C" int danet_conv_tc_set_profile_buffer(void* dev_buf) { g_tc_prof = (long long*)dev_buf; return 0; }

extern "C" int danet_conv_tc_supported(const danet_conv_desc* d) {
    tc::Geom g;
    return d && tc::make_geom(d, &g) ? 1 : 0;
}

extern "C" int64_t danet_conv_tc_packed_bytes(const danet_conv_desc* d) {
    tc::Geom g;
    if (!d || !tc::make_geom(d, &g)) return 0;
    return (int64_t)d->wsets * g.blocks_per_set * g.b_stage_bytes;
}

extern "C" int danet_conv_tc_pack(const danet_conv_desc* d, const float* w_simt, void* w_packed, danet_stream_t stream) {
    tc::Geom g;
    DANET_CHECK(d && tc::make_geom(d, &g), "danet_conv_tc_pack: shape not supported by the tcgen05 path");
    DANET_CHECK(w_simt && w_packed, "danet_conv_tc_pack: null pointer");
    const long long total = (long long)g.wsets * g.blocks_per_set * (g.b_stage_bytes / 2);
    tc::k_pack<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, w_simt, (__half*)w_packed);
    DANET_LAUNCH_CHECK();
    return 0;
}
