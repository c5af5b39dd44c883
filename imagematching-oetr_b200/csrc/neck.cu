// The OETR neck on the 5th-gen tensor cores: everything the reference runs between the
// ResNet backbone and the hot path,
//     input_proj (1x1, 1024 -> 256)                      src/model.py:45-47,118-119
//     PatchMerging: LayerNorm over channels, three stride-2 convolutions k = 4 / 8 / 16 (256 -> 256 / 128 / 128),
//     concatenated                                       src/models/backbone.py:28-67
//     input_proj2 (1x1, 512 -> 256)                      src/model.py:48-50,123-124
// as three kernels:
//   k_neck_proj  one CTA per 128 tokens: the NCHW fp32 backbone features are transposed on the fly into K-major fp16
//                operand slabs (4-slot ring), the 1x1 weights stream through a bulk-copy ring, tcgen05 accumulates
//                [128 x 256] in TMEM, the epilogue adds the bias, applies the LayerNorm (two-pass statistics) and writes
//                fp16 tokens in a PHASE-PLANAR layout xn[image][py][px][Y][X][256] (input pixel (2Y+py, 2X+px)), so that
//                every tap of a stride-2 convolution is a dense, stride-1 box of one plane.
//   k_neck_conv  the three convolutions as implicit GEMMs over the 16 x 16 tap grid of the largest kernel (the 8 x 8
//                and 4 x 4 kernels are its centre taps): per (tap, 64-channel slab) the activation operand is ONE TMA tensor
//                load (cp.async.bulk.tensor.5d, SWIZZLE_128B, zero fill outside the plane = the convolution's padding) of
//                the shifted box, the weight operand one bulk copy of the tap's pre-swizzled tile.  Items are cut by
//                convolution (see k2::Unit): k16 on tile pairs with the weights as the M operand and both activation tiles
//                as one N = 256 operand; k8 + k4 per tile.  Split-K parts per item type; every item writes its fp32
//                partial sums.
//   k_neck_out   sums the parts in fixed order (deterministic), adds the convolution biases, rounds to fp16 and runs the
//                512 -> 256 projection on the tensor cores; writes feat [n,256,h/2,w/2] fp32 NCHW (the reference's
//                feature_extraction output, and what oetr_forward reads).
// Operand precision: single fp16 operands, fp32 accumulation (measured with the oracle's rounding model: box error
// 1.1e-4 of the image side, feature error 5e-4 relative rms).  sm_100a only, no fallback.
#include "../../include/oetr_b200.h"
#include "tc_common.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <queue>
#include <vector>

using namespace oetr::tc;

namespace {

constexpr int CB = 1024;      // backbone channels (ResNet-50 layer3)
constexpr int C = 256;        // d_model
constexpr int CM = 512;       // concatenated PatchMerging channels
constexpr int TILE = 128;
constexpr float LN_EPS = 1e-5f;
constexpr uint32_t SLAB = TILE * 128;            // [128 rows x 64 fp16] = 16 KB
constexpr uint32_t WUNIT = 2 * SLAB;             // [256 rows x 64 fp16] = 32 KB
constexpr uint32_t IDESC_N128 = umma_idesc_f16(128, 128, 0, 0);
constexpr uint32_t IDESC_N256 = umma_idesc_f16(128, 256, 0, 0);
constexpr int MAX_PARTS = 16;

// ---------------------------------------------------------------------------------------------------------
// k_neck_proj
// ---------------------------------------------------------------------------------------------------------
namespace k1 {
constexpr int N_ROW = 512, N_THREADS = 576, W_PROD = 16, W_MMA = 17;
constexpr int KS = CB / 64;                      // 16 k-slabs
constexpr int A_SLOTS = 4, W_SLOTS = 3;
constexpr uint32_t SM_A = 0;
constexpr uint32_t SM_W = SM_A + A_SLOTS * SLAB;
constexpr uint32_t SM_RED = SM_W + W_SLOTS * WUNIT;             // float[2][128][4]: LayerNorm partial sums
constexpr uint32_t SM_BAR = SM_RED + 2 * 128 * 4 * 4;
constexpr uint32_t SM_TOTAL = SM_BAR + 256;
struct Bars {
    uint64_t a_full[A_SLOTS], a_free[A_SLOTS], w_full[W_SLOTS], w_empty[W_SLOTS], s_full;
    uint32_t tmem, pad;
};
static_assert(sizeof(Bars) <= 256, "Bars");
struct Params {
    const float* X;          // [n][1024][T]
    const __half* w_img;     // [16 ks][256 n][64 k] swizzled
    const float *bias, *gamma, *beta;
    __half* xn;              // [n][4][Hh][Wh][256]
    int n, h, w, T, tiles_per_img, Hh, Wh;
};
}  // namespace k1

__device__ __forceinline__ void load16(const float* __restrict__ src, size_t cstride, bool valid, float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = valid ? __ldg(src + (size_t)j * cstride) : 0.f;
}

__global__ void __launch_bounds__(k1::N_THREADS, 1) k_neck_proj(const k1::Params p) {
    using namespace k1;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = smem_u32(smem);
    const int img = blockIdx.x / p.tiles_per_img, l0 = (blockIdx.x % p.tiles_per_img) * TILE;
    if (tid == 0) {
        for (int i = 0; i < A_SLOTS; ++i) { mbar_init(&bars->a_full[i], N_ROW); mbar_init(&bars->a_free[i], 1); }
        for (int i = 0; i < W_SLOTS; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
        mbar_init(&bars->s_full, 1);
        fence_mbar_init();
    }
    if (warp == W_PROD) tmem_alloc(&bars->tmem, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem;

    if (warp == W_PROD) {
        if (lane == 0) {
            for (int ks = 0; ks < KS; ++ks) {
                const int st = ks % W_SLOTS;
                if (ks >= W_SLOTS) mbar_wait(&bars->w_empty[st], ((ks / W_SLOTS) - 1) & 1, nullptr);
                mbar_arrive_expect_tx(&bars->w_full[st], WUNIT);
                bulk_g2s(smem + SM_W + st * WUNIT, p.w_img + (size_t)ks * (WUNIT / 2), WUNIT, &bars->w_full[st]);
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        if (lane == 0) {
            for (int ks = 0; ks < KS; ++ks) {
                const int sa = ks % A_SLOTS, sw = ks % W_SLOTS;
                mbar_wait(&bars->a_full[sa], (ks / A_SLOTS) & 1, nullptr);
                mbar_wait(&bars->w_full[sw], (ks / W_SLOTS) & 1, nullptr);
                tc_fence_after();
                const uint32_t a = smem_base + SM_A + sa * SLAB, b = smem_base + SM_W + sw * WUNIT;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem, umma_desc(a + k * 32, 16, ATOM_BYTES), umma_desc(b + k * 32, 16, ATOM_BYTES), IDESC_N256,
                             (ks > 0 || k > 0) ? 1u : 0u);
                umma_commit(&bars->a_free[sa]);
                umma_commit(&bars->w_empty[sw]);
            }
            umma_commit(&bars->s_full);
        }
        __syncwarp();
    } else {
        const int q = warp & 3, cq = warp >> 2;
        const int r = q * 32 + lane;
        const int l = l0 + r;
        const bool valid = l < p.T;
        // operand slabs: this thread converts 16 channels (cq*16 ..) of token row r per k-slab; a warp's 32 lanes read 32
        // consecutive tokens of one channel (128 B, coalesced).  Loads run two slabs ahead of the stores.
        const float* src = p.X + ((size_t)img * CB + cq * 16) * p.T + (valid ? l : 0);
        const size_t T = (size_t)p.T;
        float buf[3][16];
        load16(src, T, valid, buf[0]);
        load16(src + 64 * T, T, valid, buf[1]);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            if (ks + 2 < KS) load16(src + (size_t)(ks + 2) * 64 * T, T, valid, buf[(ks + 2) % 3]);
            const int sa = ks % A_SLOTS;
            if (ks >= A_SLOTS) mbar_wait(&bars->a_free[sa], ((ks / A_SLOTS) - 1) & 1, nullptr);
            uint8_t* slab = smem + SM_A + sa * SLAB;
            *reinterpret_cast<uint4*>(slab + slab_chunk_off(r, cq * 2)) = pack8_f16(&buf[ks % 3][0]);
            *reinterpret_cast<uint4*>(slab + slab_chunk_off(r, cq * 2 + 1)) = pack8_f16(&buf[ks % 3][8]);
            fence_async_smem();
            mbar_arrive(&bars->a_full[sa]);
        }
        // epilogue: bias, LayerNorm over the 256 channels of the token (this thread: channels cq*64 .. +64)
        mbar_wait(&bars->s_full, 0, nullptr);
        tc_fence_after();
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float v0[32], v1[32];
        tmem_ld32x2_adj(tmem + lane_addr + cq * 64, v0, v1);
        tc_fence_before();
        float* red_s = reinterpret_cast<float*>(smem + SM_RED);
        float* red_m = red_s + 128 * 4;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            v0[i] += __ldg(p.bias + cq * 64 + i);
            v1[i] += __ldg(p.bias + cq * 64 + 32 + i);
            s += v0[i] + v1[i];
        }
        red_s[r * 4 + cq] = s;
        named_bar_sync(1, N_ROW);
        const float4 s4 = *reinterpret_cast<const float4*>(red_s + r * 4);
        const float mean = (s4.x + s4.y + s4.z + s4.w) * (1.f / C);
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float d0 = v0[i] - mean, d1 = v1[i] - mean;
            m2 = fmaf(d0, d0, m2);
            m2 = fmaf(d1, d1, m2);
        }
        red_m[r * 4 + cq] = m2;
        named_bar_sync(1, N_ROW);
        const float4 m4 = *reinterpret_cast<const float4*>(red_m + r * 4);
        const float rstd = rsqrtf((m4.x + m4.y + m4.z + m4.w) * (1.f / C) + LN_EPS);
        if (valid) {
            const int iy = l / p.w, ix = l - iy * p.w;
            const int ph = (iy & 1) * 2 + (ix & 1);
            __half* dst = p.xn + ((((size_t)img * 4 + ph) * p.Hh + (iy >> 1)) * p.Wh + (ix >> 1)) * C + cq * 64;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int c = cq * 64 + half * 32 + i;
                    const float x = half ? v1[i] : v0[i];
                    y[i] = fmaf((x - mean) * rstd, __ldg(p.gamma + c), __ldg(p.beta + c));
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + half * 32 + j * 8) = pack8_f16(&y[j * 8]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_PROD) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------
// k_neck_conv
// ---------------------------------------------------------------------------------------------------------
namespace k2 {
constexpr int N_THREADS = 192;                   // warp 0: TMA producer, warp 1: MMA issue, warps 2-5: epilogue
constexpr int STAGES = 4;
constexpr uint32_t A_BYTES = SLAB, STAGE = 3 * SLAB;              // 48 KB: {A0, A1, B 16 KB} or {A, B 32 KB}
constexpr uint32_t SM_BAR = STAGES * STAGE;
constexpr uint32_t SM_TOTAL = SM_BAR + 128;
struct Bars {
    uint64_t full[STAGES], empty[STAGES], done;
    uint32_t tmem, pad;
};
// One unit = one (tap, 64-channel slab) step of an item.  The L2 -> SM operand stream bounds this kernel (a 16 KB A tile and
// a 16 KB weight tile for 256 MMA cycles = 124 B/cycle against ~64 B/cycle of ingest), so the work is cut so that every
// weight tile is used twice:
//   kind 0  k16 (16 x 16 taps, 128 outputs): an item covers a PAIR of tiles; per unit two A tiles and ONE 128-row weight
//           tile, two N=128 MMAs into TMEM columns [0,128) and [128,256)  (46 KB per 512 MMA cycles = 92 B/cycle)
//   kind 1  k8 (the 8 x 8 centre taps, 128 outputs): one tile, 128-row weight tile, columns [0,128)
//   kind 2  k4 (the 4 x 4 centre taps, 256 outputs): one tile, 256-row weight tile, columns [128,384)
// Type-1 items hold kind-0 units, type-2 items kind-1 and kind-2 units; both are split over K into parts.
struct Unit { int dx, dy, phase_slab_kind; uint32_t w_kb; };
struct Params {
    const Unit* units1;      // kind-0 units in the order of the chosen part count P1 (part p: [begin1[p], begin1[p+1]))
    const Unit* units2;      // kind-1/2 units for P2
    const __half* w_img;
    float* partial1;         // [P1][tiles_pad][128][128]
    float* partial2;         // [P2][tiles][128][384]
    int begin1[MAX_PARTS + 1], begin2[MAX_PARTS + 1];
    int P1, P2, pairs, tiles, YG, ny, nb, R, items1;
};
}  // namespace k2

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(k2::N_THREADS, 1) k_neck_conv(const __grid_constant__ CUtensorMap tmap, const k2::Params p) {
    using namespace k2;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = smem_u32(smem);
    // items: type 1 first (the long ones), part-major so that neighbouring CTAs stream the same weights
    const bool type1 = (int)blockIdx.x < p.items1;
    const int local = type1 ? (int)blockIdx.x : (int)blockIdx.x - p.items1;
    const int part = type1 ? local / p.pairs : local / p.tiles;
    const int tile0 = type1 ? 2 * (local % p.pairs) : local % p.tiles;
    const Unit* units = type1 ? p.units1 : p.units2;
    const int ub = type1 ? p.begin1[part] : p.begin2[part], ue = type1 ? p.begin1[part + 1] : p.begin2[part + 1];
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        mbar_init(&bars->done, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&bars->tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem;

    if (warp == 0) {
        if (lane == 0) {
            const int y0 = (tile0 % p.YG) * p.ny, n0 = (tile0 / p.YG) * p.nb;
            const int y1 = ((tile0 + 1) % p.YG) * p.ny, n1 = ((tile0 + 1) / p.YG) * p.nb;      // second tile of a pair (may lie
            const uint32_t a_bytes = (uint32_t)p.R * 128u;                                     // past the last image: zero fill)
            for (int u = ub; u < ue; ++u) {
                const int i = u - ub, st = i % STAGES;
                if (i >= STAGES) mbar_wait(&bars->empty[st], ((i / STAGES) - 1) & 1, nullptr);
                const Unit un = units[u];
                const int phase = un.phase_slab_kind & 0xff, slab = (un.phase_slab_kind >> 8) & 0xff, kind = un.phase_slab_kind >> 16;
                const uint32_t b_bytes = kind == 2 ? WUNIT : SLAB;
                const uint32_t stage = smem_base + st * STAGE;
                mbar_arrive_expect_tx(&bars->full[st], (kind == 0 ? 2 * a_bytes : a_bytes) + b_bytes);
                tma_load_5d(stage, &tmap, slab * 64, un.dx, y0 + un.dy, phase, n0, &bars->full[st]);
                if (kind == 0) tma_load_5d(stage + A_BYTES, &tmap, slab * 64, un.dx, y1 + un.dy, phase, n1, &bars->full[st]);
                bulk_g2s(smem + st * STAGE + (kind == 0 ? 2 * A_BYTES : A_BYTES), p.w_img + (size_t)un.w_kb * 512, b_bytes, &bars->full[st]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            bool init0 = false, init1 = false;       // type 1: both tile accumulators together; type 2: k8 | k4
            for (int u = ub; u < ue; ++u) {
                const int i = u - ub, st = i % STAGES;
                const int kind = units[u].phase_slab_kind >> 16;
                mbar_wait(&bars->full[st], (i / STAGES) & 1, nullptr);
                tc_fence_after();
                const uint32_t a = smem_base + st * STAGE;
                if (kind == 0) {
                    // swapped roles: the WEIGHT tile is the M operand (128 output channels = TMEM lanes), the two
                    // activation tiles, adjacent in the stage, are ONE N = 256 operand (positions = TMEM columns):
                    // an N=256 MMA fetches 12 KB of operands per 128 cycles, two N=128 MMAs 16 KB -- this kernel is bound
                    // by shared-memory bandwidth (TMA writes + MMA operand fetch), measured
                    const uint32_t w = a + 2 * A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(tmem, umma_desc(w + k * 32, 16, ATOM_BYTES), umma_desc(a + k * 32, 16, ATOM_BYTES), IDESC_N256,
                                 (!init0 && k == 0) ? 0u : 1u);
                    init0 = true;
                } else {
                    const uint32_t b = a + A_BYTES;
                    const bool first = kind == 1 ? !init0 : !init1;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(tmem + (kind == 2 ? 128u : 0u), umma_desc(a + k * 32, 16, ATOM_BYTES), umma_desc(b + k * 32, 16, ATOM_BYTES),
                                 kind == 2 ? IDESC_N256 : IDESC_N128, (first && k == 0) ? 0u : 1u);
                    if (kind == 1) init0 = true; else init1 = true;
                }
                umma_commit(&bars->empty[st]);
            }
            umma_commit(&bars->done);
        }
        __syncwarp();
    } else {
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        mbar_wait(&bars->done, 0, nullptr);
        tc_fence_after();
        if (type1) {                                 // TMEM lane = output channel, column = position: [part][tile][pos][128 oc]
            const int oc = r;
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                float* out = p.partial1 + ((size_t)part * 2 * p.pairs + tile0 + t) * TILE * 128 + oc;
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    float v[32];
                    tmem_ld32(tmem + lane_addr + t * 128 + cc * 32, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (cc * 32 + j < p.R) out[(size_t)(cc * 32 + j) * 128] = v[j];          // a warp: 32 channels of one position
                }
            }
        } else {                                     // [part][tile][128][384]: k8 | k4
            float* out = p.partial2 + (((size_t)part * p.tiles + tile0) * TILE + r) * 384;
#pragma unroll 1
            for (int cc = 0; cc < 12; ++cc) {
                float v[32];
                tmem_ld32(tmem + lane_addr + cc * 32, v);
                if (r < p.R) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(out + cc * 32 + j * 4) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------
// k_neck_out
// ---------------------------------------------------------------------------------------------------------
namespace k3 {
constexpr int N_ROW = 512, N_THREADS = 576, W_PROD = 16, W_MMA = 17;
constexpr int KS = CM / 64;                      // 8
constexpr int W_SLOTS = 2;
constexpr uint32_t SM_A = 0;                     // [128 x 512] fp16 operand image: 8 slabs
constexpr uint32_t SM_W = SM_A + KS * SLAB;
constexpr uint32_t SM_BAR = SM_W + W_SLOTS * WUNIT;
constexpr uint32_t SM_TOTAL = SM_BAR + 128;
struct Bars {
    uint64_t a_full, w_full[W_SLOTS], w_empty[W_SLOTS], s_full;
    uint32_t tmem, pad;
};
struct Params {
    const float* partial1;   // [P1][tiles_pad][128][128]: k16
    const float* partial2;   // [P2][tiles][128][384]: k8 | k4
    const __half* w_img;     // [8 ks][256 n][64 k] swizzled, K in the order k16 | k8 | k4
    const float *bias_cat, *bias2;       // [512] in the same order; [256]
    float* feat;             // [n][256][ho][wo]
    int P1, P2, tiles_pad, tiles, YG, ny, nb, R, n, ho, wo;
};
}  // namespace k3

__global__ void __launch_bounds__(k3::N_THREADS, 1) k_neck_out(const k3::Params p) {
    using namespace k3;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = smem_u32(smem);
    const int tile = blockIdx.x;
    if (tid == 0) {
        mbar_init(&bars->a_full, N_ROW);
        for (int i = 0; i < W_SLOTS; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
        mbar_init(&bars->s_full, 1);
        fence_mbar_init();
    }
    if (warp == W_PROD) tmem_alloc(&bars->tmem, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem;

    if (warp == W_PROD) {
        if (lane == 0) {
            for (int ks = 0; ks < KS; ++ks) {
                const int st = ks % W_SLOTS;
                if (ks >= W_SLOTS) mbar_wait(&bars->w_empty[st], ((ks / W_SLOTS) - 1) & 1, nullptr);
                mbar_arrive_expect_tx(&bars->w_full[st], WUNIT);
                bulk_g2s(smem + SM_W + st * WUNIT, p.w_img + (size_t)ks * (WUNIT / 2), WUNIT, &bars->w_full[st]);
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        if (lane == 0) {
            mbar_wait(&bars->a_full, 0, nullptr);
            for (int ks = 0; ks < KS; ++ks) {
                const int st = ks % W_SLOTS;
                mbar_wait(&bars->w_full[st], (ks / W_SLOTS) & 1, nullptr);
                tc_fence_after();
                const uint32_t a = smem_base + SM_A + ks * SLAB, b = smem_base + SM_W + st * WUNIT;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem, umma_desc(a + k * 32, 16, ATOM_BYTES), umma_desc(b + k * 32, 16, ATOM_BYTES), IDESC_N256,
                             (ks > 0 || k > 0) ? 1u : 0u);
                umma_commit(&bars->w_empty[st]);
            }
            umma_commit(&bars->s_full);
        }
        __syncwarp();
    } else {
        const int q = warp & 3, cq = warp >> 2;
        const int r = q * 32 + lane;
        const bool in_box = r < p.R;
        // concatenated activations of the row: sum of the parts (fixed order) + convolution biases, 128 columns per thread
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            const int c0 = cq * 128 + ch * 32;
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias_cat + c0) + j);
                v[4 * j] = b4.x; v[4 * j + 1] = b4.y; v[4 * j + 2] = b4.z; v[4 * j + 3] = b4.w;
            }
            if (in_box) {
                // columns [0,128) = k16 from the type-1 parts, [128,512) = k8 | k4 from the type-2 parts
                const int nparts = cq == 0 ? p.P1 : p.P2;
                for (int part = 0; part < nparts; ++part) {
                    const float4* src = cq == 0
                        ? reinterpret_cast<const float4*>(p.partial1 + (((size_t)part * p.tiles_pad + tile) * TILE + r) * 128 + ch * 32)
                        : reinterpret_cast<const float4*>(p.partial2 + (((size_t)part * p.tiles + tile) * TILE + r) * 384 + (c0 - 128));
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = __ldg(src + j);
                        v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                    }
                }
            }
            store_row32_f16(smem + SM_A, SLAB, r, c0, v);
        }
        fence_async_smem();
        mbar_arrive(&bars->a_full);
        mbar_wait(&bars->s_full, 0, nullptr);
        tc_fence_after();
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float v0[32], v1[32];
        tmem_ld32x2_adj(tmem + lane_addr + cq * 64, v0, v1);
        // row r of the tile = box element (x, yy, nn), x fastest
        const int x = r % p.wo, t1 = r / p.wo, yy = t1 % p.ny, nn = t1 / p.ny;
        const int oy = (tile % p.YG) * p.ny + yy, img = (tile / p.YG) * p.nb + nn;
        if (in_box && oy < p.ho && img < p.n) {
            const size_t plane = (size_t)p.ho * p.wo;
            float* dst = p.feat + ((size_t)img * C + cq * 64) * plane + (size_t)oy * p.wo + x;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                dst[(size_t)i * plane] = v0[i] + __ldg(p.bias2 + cq * 64 + i);
                dst[(size_t)(32 + i) * plane] = v1[i] + __ldg(p.bias2 + cq * 64 + 32 + i);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_PROD) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
thread_local char g_nerr[512] = "";
int nfail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_nerr, sizeof(g_nerr), fmt, ap);
    va_end(ap);
    return code;
}
#define NCU(call)                                                                                            \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return nfail(e_ == cudaErrorMemoryAllocation ? OETR_E_NOMEM : OETR_E_CUDA, "%s: %s (%s:%d)", #call, \
                         cudaGetErrorString(e_), __FILE__, __LINE__);                                        \
    } while (0)

// offsets (floats) of the packed neck weights, NECK_ORDER of oetr_b200/weights.py
struct NeckLayout {
    size_t w1, b1, ln_g, ln_b, w4, b4, w8, b8, w16, b16, w2, b2, total;
};
NeckLayout neck_layout() {
    NeckLayout L{};
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += n; return r; };
    L.w1 = take((size_t)C * CB); L.b1 = take(C);
    L.ln_g = take(C); L.ln_b = take(C);
    L.w4 = take((size_t)256 * C * 16); L.b4 = take(256);
    L.w8 = take((size_t)128 * C * 64); L.b8 = take(128);
    L.w16 = take((size_t)128 * C * 256); L.b16 = take(128);
    L.w2 = take((size_t)C * CM); L.b2 = take(C);
    L.total = o;
    return L;
}

// element (row n, column k) of a swizzled [rows x 64] fp16 tile
inline size_t tile_half_index(int n, int k) { return (slab_chunk_off((uint32_t)n, (uint32_t)(k >> 3)) >> 1) + (k & 7); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Geometry {
    int n, h, w, T, tiles_per_img, Hh, Wh, ho, wo, ny, nb, YG, NG, tiles, pairs, R, P1, P2;
    size_t xn_bytes, partial1_bytes, partial2_bytes;
};

// makespan (cycles) of n1 items of cost c1 followed by n2 items of cost c2 dispatched in order to `sms` SMs
double makespan(long n1, double c1, long n2, double c2, int sms) {
    std::priority_queue<double, std::vector<double>, std::greater<double>> q;
    for (int i = 0; i < sms; ++i) q.push(0.0);
    double end = 0.0;
    auto run = [&](long n, double c) {
        for (long i = 0; i < n; ++i) { const double t = q.top() + c; q.pop(); q.push(t); if (t > end) end = t; }
    };
    run(n1, c1);
    run(n2, c2);
    return end;
}

// box rows of a conv tile = wo * ny * nb <= 128: the (ny, nb) with the fewest tiles; then the split-K part counts of the two
// item types (k16 on tile pairs; k8 + k4 per tile) that minimise the modelled makespan on `sms` SMs plus the round trip of
// the partial sums through HBM.  Step costs (cycles) follow the measured L2 -> SM ingest of ~64 B/cycle per SM.
int make_geometry(int n, int h, int w, int sms, Geometry& g) {
    if (n < 1 || h < 2 || w < 2 || h > 200 || w > 200) return -1;
    g.n = n; g.h = h; g.w = w; g.T = h * w;
    g.tiles_per_img = (g.T + TILE - 1) / TILE;
    g.Hh = (h + 1) / 2; g.Wh = (w + 1) / 2; g.ho = h / 2; g.wo = w / 2;
    long best = -1;
    for (int ny = 1; ny <= g.ho && g.wo * ny <= TILE; ++ny)
        for (int nb = 1; nb <= n && g.wo * ny * nb <= TILE; ++nb) {
            const long tiles = (long)((g.ho + ny - 1) / ny) * ((n + nb - 1) / nb);
            if (best < 0 || tiles < best || (tiles == best && ny > g.ny)) { best = tiles; g.ny = ny; g.nb = nb; }
        }
    g.YG = (g.ho + g.ny - 1) / g.ny; g.NG = (n + g.nb - 1) / g.nb;
    g.tiles = g.YG * g.NG; g.R = g.wo * g.ny * g.nb;
    g.pairs = (g.tiles + 1) / 2;
    const double step16 = 1000.0, step8 = 500.0, step4 = 750.0, fixed = 40000.0, hbm_bytes_per_cycle = 2000.0;   // fitted to a sweep on B200
    double best_cost = -1.0;
    static const int force1 = getenv("OETR_NECK_P1") ? atoi(getenv("OETR_NECK_P1")) : 0;      // experiments
    static const int force2 = getenv("OETR_NECK_P2") ? atoi(getenv("OETR_NECK_P2")) : 0;
    for (int P1 : {1, 2, 4, 8, 16})
        for (int P2 : {1, 2, 4, 8}) {
            if ((force1 && P1 != force1) || (force2 && P2 != force2)) continue;
            const double c1 = 1024.0 / P1 * step16 + fixed, c2 = (256.0 * step8 + 64.0 * step4) / P2 + fixed;
            const double bytes = 2.0 * ((double)P1 * 2 * g.pairs * TILE * 128 * 4 + (double)P2 * g.tiles * TILE * 384 * 4);
            const double cost = makespan((long)g.pairs * P1, c1, (long)g.tiles * P2, c2, sms) + bytes / hbm_bytes_per_cycle;
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; g.P1 = P1; g.P2 = P2; }
        }
    g.xn_bytes = (((size_t)n * 4 * g.Hh * g.Wh * C * sizeof(__half)) + 1023) & ~(size_t)1023;
    g.partial1_bytes = (size_t)g.P1 * 2 * g.pairs * TILE * 128 * sizeof(float);
    g.partial2_bytes = (size_t)g.P2 * g.tiles * TILE * 384 * sizeof(float);
    return 0;
}

std::mutex g_attr_mu;
bool g_attr_done[64] = {};

}  // namespace

struct oetr_neck {
    int device = 0, sms = 148;
    __half *w1_img = nullptr, *conv_img = nullptr, *w2_img = nullptr;
    float* vec = nullptr;            // b1[256] | ln_g[256] | ln_b[256] | bias_cat[512] | b2[256]
    k2::Unit *units1 = nullptr, *units2 = nullptr;      // unit lists of every part count: [MAX_PARTS + 1][n_units]
    int n_units1 = 0, n_units2 = 0;
    int begin1[MAX_PARTS + 1][MAX_PARTS + 1] = {}, begin2[MAX_PARTS + 1][MAX_PARTS + 1] = {};
    EncodeTiledFn encode = nullptr;
    int last_launches = 0;
    std::mutex mu;                   // geometry cache (the part-count search costs ~0.1 ms: once per problem size)
    std::vector<Geometry> cache;
};

namespace {
int cached_geometry(oetr_neck* h, int n, int height, int width, Geometry& g) {
    std::lock_guard<std::mutex> lk(h->mu);
    for (const Geometry& c : h->cache)
        if (c.n == n && c.h == height && c.w == width) { g = c; return 0; }
    if (make_geometry(n, height, width, h->sms, g)) return -1;
    if (h->cache.size() >= 64) h->cache.erase(h->cache.begin());
    h->cache.push_back(g);
    return 0;
}
}  // namespace

extern "C" {

const char* oetr_neck_last_error(void) { return g_nerr; }
size_t oetr_neck_packed_weight_count(void) { return neck_layout().total; }

int oetr_neck_destroy(oetr_neck* h) {
    if (!h) return OETR_OK;
    cudaFree(h->w1_img); cudaFree(h->conv_img); cudaFree(h->w2_img); cudaFree(h->vec); cudaFree(h->units1); cudaFree(h->units2);
    delete h;
    return OETR_OK;
}

int oetr_neck_create(const float* weights_host, size_t n_floats, oetr_neck** out) {
    if (!weights_host || !out) return nfail(OETR_E_ARG, "oetr_neck_create: null argument");
    *out = nullptr;
    const NeckLayout L = neck_layout();
    if (n_floats != L.total) return nfail(OETR_E_ARG, "oetr_neck_create: %zu weights given, %zu expected", n_floats, L.total);
    int dev = 0;
    NCU(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    NCU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return nfail(OETR_E_ARCH, "oetr_neck_create: device %d is sm_%d%d, this library is sm_100a only", dev, prop.major, prop.minor);
    oetr_neck* h = new (std::nothrow) oetr_neck();
    if (!h) return nfail(OETR_E_NOMEM, "oetr_neck_create: host allocation failed");
    h->device = dev; h->sms = prop.multiProcessorCount;
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            delete h;
            return nfail(OETR_E_CUDA, "oetr_neck_create: cuTensorMapEncodeTiled is not available from the driver");
        }
        h->encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const float* W = weights_host;
    // input_proj: 16 k-slabs of [256 n x 64 k]
    std::vector<__half> w1((size_t)k1::KS * 256 * 64);
    for (int ks = 0; ks < k1::KS; ++ks)
        for (int n = 0; n < C; ++n)
            for (int k = 0; k < 64; ++k)
                w1[(size_t)ks * 256 * 64 + tile_half_index(n, k)] = __float2half_rn(W[L.w1 + (size_t)n * CB + ks * 64 + k]);
    // input_proj2 with K in the TMEM column order: k16 (concat 384..511) | k8 (256..383) | k4 (0..255)
    auto cat_of = [](int j) { return j < 128 ? 384 + j : (j < 256 ? 256 + (j - 128) : j - 256); };
    std::vector<__half> w2((size_t)k3::KS * 256 * 64);
    for (int ks = 0; ks < k3::KS; ++ks)
        for (int n = 0; n < C; ++n)
            for (int k = 0; k < 64; ++k)
                w2[(size_t)ks * 256 * 64 + tile_half_index(n, k)] = __float2half_rn(W[L.w2 + (size_t)n * CM + cat_of(ks * 64 + k)]);
    std::vector<float> vec(256 * 3 + 512 + 256);
    memcpy(&vec[0], W + L.b1, 256 * 4); memcpy(&vec[256], W + L.ln_g, 256 * 4); memcpy(&vec[512], W + L.ln_b, 256 * 4);
    memcpy(&vec[768], W + L.b16, 128 * 4); memcpy(&vec[768 + 128], W + L.b8, 128 * 4); memcpy(&vec[768 + 256], W + L.b4, 256 * 4);
    memcpy(&vec[768 + 512], W + L.b2, 256 * 4);
    // convolution units; every unit's weight tile starts on a 1 KB boundary.  kind 0: k16 over all 256 taps (128 rows);
    // kind 1: k8 over its 64 taps (128 rows); kind 2: k4 over its 16 taps (256 rows)
    std::vector<k2::Unit> u16, u8, u4;
    size_t kb = 0;
    for (int kind = 0; kind < 3; ++kind)
        for (int ky = 0; ky < 16; ++ky)
            for (int kx = 0; kx < 16; ++kx) {
                const bool in8 = ky >= 4 && ky < 12 && kx >= 4 && kx < 12, in4 = ky >= 6 && ky < 10 && kx >= 6 && kx < 10;
                if ((kind == 1 && !in8) || (kind == 2 && !in4)) continue;
                const int py = (ky - 7) & 1, px = (kx - 7) & 1;
                for (int slab = 0; slab < 4; ++slab) {
                    k2::Unit u;
                    u.dy = (ky - 7 - py) / 2; u.dx = (kx - 7 - px) / 2;
                    u.phase_slab_kind = (py * 2 + px) | (slab << 8) | (kind << 16);
                    u.w_kb = (uint32_t)kb;
                    kb += kind == 2 ? 32 : 16;
                    (kind == 0 ? u16 : (kind == 1 ? u8 : u4)).push_back(u);
                }
            }
    std::vector<__half> conv(kb * 512);
    for (auto* lst : {&u16, &u8, &u4})
        for (const k2::Unit& u : *lst) {
            const int py = (u.phase_slab_kind & 0xff) >> 1, px = u.phase_slab_kind & 1;
            const int ky = 2 * u.dy + py + 7, kx = 2 * u.dx + px + 7;
            const int slab = (u.phase_slab_kind >> 8) & 0xff, kind = u.phase_slab_kind >> 16;
            __half* t = conv.data() + (size_t)u.w_kb * 512;
            const int rows = kind == 2 ? 256 : 128;
            for (int n = 0; n < rows; ++n)
                for (int k = 0; k < 64; ++k) {
                    const int c = slab * 64 + k;
                    float v;
                    if (kind == 2) v = W[L.w4 + (((size_t)n * C + c) * 4 + (ky - 6)) * 4 + (kx - 6)];
                    else if (kind == 0) v = W[L.w16 + (((size_t)n * C + c) * 16 + ky) * 16 + kx];
                    else v = W[L.w8 + (((size_t)n * C + c) * 8 + (ky - 4)) * 8 + (kx - 4)];
                    t[tile_half_index(n, k)] = __float2half_rn(v);
                }
        }
    h->n_units1 = (int)u16.size();
    h->n_units2 = (int)(u8.size() + u4.size());
    // the unit list of every part count P: part p takes every P-th unit of each kind (type 2: k8 units first, then k4)
    std::vector<k2::Unit> all1((size_t)(MAX_PARTS + 1) * h->n_units1), all2((size_t)(MAX_PARTS + 1) * h->n_units2);
    for (int P = 1; P <= MAX_PARTS; ++P) {
        size_t o1 = (size_t)P * h->n_units1, o2 = (size_t)P * h->n_units2;
        const size_t b1 = o1, b2 = o2;
        for (int part = 0; part < P; ++part) {
            h->begin1[P][part] = (int)(o1 - b1);
            h->begin2[P][part] = (int)(o2 - b2);
            for (size_t i = part; i < u16.size(); i += P) all1[o1++] = u16[i];
            for (size_t i = part; i < u8.size(); i += P) all2[o2++] = u8[i];
            for (size_t i = part; i < u4.size(); i += P) all2[o2++] = u4[i];
        }
        h->begin1[P][P] = (int)(o1 - b1);
        h->begin2[P][P] = (int)(o2 - b2);
    }
    cudaError_t e = cudaSuccess;
    auto up = [&](void** d, const void* src, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(d, bytes);
        if (e == cudaSuccess) e = cudaMemcpy(*d, src, bytes, cudaMemcpyHostToDevice);
    };
    up((void**)&h->w1_img, w1.data(), w1.size() * 2);
    up((void**)&h->w2_img, w2.data(), w2.size() * 2);
    up((void**)&h->conv_img, conv.data(), conv.size() * 2);
    up((void**)&h->vec, vec.data(), vec.size() * 4);
    up((void**)&h->units1, all1.data(), all1.size() * sizeof(k2::Unit));
    up((void**)&h->units2, all2.data(), all2.size() * sizeof(k2::Unit));
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_attr_mu);
        if (dev < 64 && !g_attr_done[dev]) {
            e = cudaFuncSetAttribute(k_neck_proj, cudaFuncAttributeMaxDynamicSharedMemorySize, k1::SM_TOTAL);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_neck_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, k2::SM_TOTAL);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_neck_out, cudaFuncAttributeMaxDynamicSharedMemorySize, k3::SM_TOTAL);
            if (e == cudaSuccess) g_attr_done[dev] = true;
        }
    }
    if (e != cudaSuccess) {
        oetr_neck_destroy(h);
        return nfail(e == cudaErrorMemoryAllocation ? OETR_E_NOMEM : OETR_E_CUDA, "oetr_neck_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return OETR_OK;
}

int oetr_neck_workspace_bytes(const oetr_neck* h, int n_images, int height, int width, size_t* out) {
    if (!h || !out) return nfail(OETR_E_ARG, "oetr_neck_workspace_bytes: null argument");
    Geometry g;
    if (cached_geometry(const_cast<oetr_neck*>(h), n_images, height, width, g)) return nfail(OETR_E_SHAPE, "oetr_neck_workspace_bytes: %d images of %d x %d", n_images, height, width);
    *out = g.xn_bytes + g.partial1_bytes + g.partial2_bytes + 1024;
    return OETR_OK;
}

int oetr_neck_last_launch_count(const oetr_neck* h) { return h ? h->last_launches : 0; }

// host-only: the conv tiling chosen for a problem (tests): out = {tiles, rows per tile, ny, nb, parts of the k16 items
// (on tile pairs) * 100 + parts of the k8 + k4 items}
int oetr_neck_geometry(int n_images, int height, int width, int sms, int* out5) {
    Geometry g;
    if (!out5 || make_geometry(n_images, height, width, sms > 0 ? sms : 148, g)) return nfail(OETR_E_SHAPE, "oetr_neck_geometry: bad problem");
    out5[0] = g.tiles; out5[1] = g.R; out5[2] = g.ny; out5[3] = g.nb; out5[4] = g.P1 * 100 + g.P2;
    return OETR_OK;
}

int oetr_neck_forward(oetr_neck* h, const float* backbone_out, int n_images, int height, int width, float* feat_out,
                      void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !backbone_out || !feat_out || !workspace) return nfail(OETR_E_ARG, "oetr_neck_forward: null argument");
    int dev = -1;
    NCU(cudaGetDevice(&dev));
    if (dev != h->device) return nfail(OETR_E_ARG, "oetr_neck_forward: handle belongs to device %d, current device is %d", h->device, dev);
    Geometry g;
    if (cached_geometry(h, n_images, height, width, g)) return nfail(OETR_E_SHAPE, "oetr_neck_forward: %d images of %d x %d", n_images, height, width);
    const uintptr_t base = ((uintptr_t)workspace + 1023) & ~(uintptr_t)1023;
    if (base + g.xn_bytes + g.partial1_bytes + g.partial2_bytes > (uintptr_t)workspace + workspace_bytes)
        return nfail(OETR_E_NOMEM, "oetr_neck_forward: workspace of %zu bytes, %zu needed", workspace_bytes,
                     g.xn_bytes + g.partial1_bytes + g.partial2_bytes + 1024);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    __half* xn = reinterpret_cast<__half*>(base);
    float* partial1 = reinterpret_cast<float*>(base + g.xn_bytes);
    float* partial2 = reinterpret_cast<float*>(base + g.xn_bytes + g.partial1_bytes);
    int launches = 0;
    if ((height & 1) || (width & 1)) NCU(cudaMemsetAsync(xn, 0, g.xn_bytes, s));     // the missing last row / column of the odd planes
    {
        k1::Params p;
        p.X = backbone_out; p.w_img = h->w1_img; p.bias = h->vec; p.gamma = h->vec + 256; p.beta = h->vec + 512; p.xn = xn;
        p.n = g.n; p.h = g.h; p.w = g.w; p.T = g.T; p.tiles_per_img = g.tiles_per_img; p.Hh = g.Hh; p.Wh = g.Wh;
        k_neck_proj<<<g.n * g.tiles_per_img, k1::N_THREADS, k1::SM_TOTAL, s>>>(p);
        NCU(cudaGetLastError());
        ++launches;
    }
    {
        CUtensorMap tmap;
        const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)g.Wh, (cuuint64_t)g.Hh, 4, (cuuint64_t)g.n};
        const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * g.Wh, (cuuint64_t)C * 2 * g.Wh * g.Hh,
                                       (cuuint64_t)C * 2 * g.Wh * g.Hh * 4};
        const cuuint32_t box[5] = {64, (cuuint32_t)g.wo, (cuuint32_t)g.ny, 1, (cuuint32_t)g.nb};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        const CUresult r = h->encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, xn, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return nfail(OETR_E_CUDA, "oetr_neck_forward: cuTensorMapEncodeTiled failed (%d)", (int)r);
        k2::Params p;
        p.units1 = h->units1 + (size_t)g.P1 * h->n_units1; p.units2 = h->units2 + (size_t)g.P2 * h->n_units2;
        p.w_img = h->conv_img; p.partial1 = partial1; p.partial2 = partial2;
        for (int i = 0; i <= MAX_PARTS; ++i) {
            p.begin1[i] = h->begin1[g.P1][i < g.P1 ? i : g.P1];
            p.begin2[i] = h->begin2[g.P2][i < g.P2 ? i : g.P2];
        }
        p.P1 = g.P1; p.P2 = g.P2; p.pairs = g.pairs; p.tiles = g.tiles; p.YG = g.YG; p.ny = g.ny; p.nb = g.nb; p.R = g.R;
        p.items1 = g.pairs * g.P1;
        k_neck_conv<<<p.items1 + g.tiles * g.P2, k2::N_THREADS, k2::SM_TOTAL, s>>>(tmap, p);
        NCU(cudaGetLastError());
        ++launches;
    }
    {
        k3::Params p;
        p.partial1 = partial1; p.partial2 = partial2; p.w_img = h->w2_img; p.bias_cat = h->vec + 768; p.bias2 = h->vec + 768 + 512; p.feat = feat_out;
        p.P1 = g.P1; p.P2 = g.P2; p.tiles_pad = 2 * g.pairs; p.tiles = g.tiles; p.YG = g.YG; p.ny = g.ny; p.nb = g.nb; p.R = g.R; p.n = g.n; p.ho = g.ho; p.wo = g.wo;
        k_neck_out<<<g.tiles, k3::N_THREADS, k3::SM_TOTAL, s>>>(p);
        NCU(cudaGetLastError());
        ++launches;
    }
    h->last_launches = launches;
    return OETR_OK;
}

}  // extern "C"
