// Shared device code of the tcgen05 kernels: tile constants, shared-memory map, barriers, the fp32 -> (hi, lo) fp16
// operand-image stores, tile geometry (per-image and flat), special functions, and the producer / MMA-issue building blocks.
// Part of the tcgen05 (OETR_PREC_FP16) path; compiled into tc_kernels.cu (one translation unit: kernels are
// launched from the host code there).
#pragma once
#include "tc_common.cuh"
#include "tc_path.cuh"

namespace oetr {
using namespace tc;


constexpr int TILE = 128;                       // tokens per CTA
constexpr int N_ROW_THREADS = 512;              // warps 0-15
constexpr int WARP_PRODUCER = 16, WARP_MMA = 17;
constexpr int N_THREADS = 576;
constexpr int STAGE_HALFS = 128 * 64;           // one ring stage: [128 N rows][64 K cols] fp16, swizzled, 16 KB
constexpr uint32_t STAGE_BYTES = STAGE_HALFS * 2;
constexpr int RING = 6;                         // 3 units of two adjacent stages (one [256 x 64] B tile each)
constexpr int GEMM_STAGES = 16;                 // a 256x256 weight block: 4 k-slabs x {hi n0, hi n1, lo n0, lo n1}
constexpr size_t GEMM_HALFS = (size_t)GEMM_STAGES * STAGE_HALFS;      // 256 KB
constexpr uint32_t SLAB_BYTES = TILE * 128;     // one [128 x 64] fp16 operand slab = 16 KB
constexpr uint32_t IMG_BYTES = 4 * SLAB_BYTES;  // a [128 x 256] fp16 operand image = 64 KB
// per encoder layer: Wq | W1a | W1b | W2a | W2b | Wv | Wk | Wm (the last only for full attention: linear attention folds
// the merge projection into per-image weights, k_fold) ; per decoder layer: Wv | Wk
constexpr int ENC_LAYER_GEMMS = 8, DEC_LAYER_GEMMS = 2;
constexpr size_t ENC_LAYER_HALFS = ENC_LAYER_GEMMS * GEMM_HALFS;
constexpr size_t DEC_LAYER_HALFS = DEC_LAYER_GEMMS * GEMM_HALFS;
constexpr size_t DEC_T_FLOATS = (size_t)6 * C * C + (size_t)2 * FF * C;   // transposed fp32 decoder weights per layer

// one partial linear-attention summary of a (tile, image): KV[8][32][32] | Ksum of the four row quarters [4][256]
constexpr int PART_FLOATS = NH * HD * HD + 4 * C;

constexpr uint32_t IDESC_N256 = umma_idesc_f16(128, 256, 0, 0);
constexpr uint32_t IDESC_KV = umma_idesc_f16(128, 128, 1, 1);      // both operands MN-major (token = K)

// shared-memory map (dynamic, 1024-byte aligned)
constexpr uint32_t SM_AHI = 0;                                     // operand image, hi part (64 KB)
constexpr uint32_t SM_ALO = SM_AHI + IMG_BYTES;                    // operand image, lo part (64 KB)
constexpr uint32_t SM_RING = SM_ALO + IMG_BYTES;                   // weight ring: RING stages of 16 KB (3 units of 32 KB)
constexpr uint32_t SM_X = SM_RING + RING * STAGE_BYTES;            // float[512] scratch of the row warps: LayerNorm
                                                                   // partials -> (gamma | beta) -> Ksum of the source
                                                                   // image -> Ksum exchange (one user at a time)
constexpr uint32_t SM_BAR = SM_X + 512 * 4;                        // mbarriers + tmem pointer
constexpr uint32_t SM_TOTAL = SM_BAR + 256;
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");
static_assert(sizeof(uint64_t) * (2 * RING + 8) + 8 <= 256, "Bars must fit its reservation");
// the kv phase re-uses the operand image space for the MN-major half images (tokens = K dimension)
constexpr uint32_t KF_OFF = 0;                                     // Kf half image: 2 slabs (32 KB) inside hi / lo
constexpr uint32_t V_OFF = 2 * SLAB_BYTES;                         // V  half image: 2 slabs (32 KB) inside hi / lo

// OETR_TIMING=1: global cycle accumulators of the k_enc launches with a query and a source phase (atomicAdd per CTA)
// DBG_LOG: optional CTA log (OETR_TIMING=2): slot DBG_LOG_N counts entries of {smid, kernel id, t_start ns, t_end ns} appended after the
// accumulators (tools/sm_timeline.py reconstructs per-SM busy time from it)
enum { DBG_MMA_TOTAL = 0, DBG_MMA_WAIT_A = 1, DBG_MMA_WAIT_W = 2, DBG_TILES = 3, DBG_TILE_NS = 4, DBG_STAGE0 = 8, DBG_CONV = 40, DBG_LOG_N = 47, DBG_SLOTS = 48, DBG_LOG_CAP = 1 << 18 };
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t sm_id() { uint32_t s; asm volatile("mov.u32 %0, %smid;" : "=r"(s)); return s; }
__device__ __forceinline__ void dbg_log_cta(unsigned long long* acc, int kernel_id, unsigned long long t0) {
    if (!acc || !acc[DBG_LOG_N + 1]) return;               // slot 48 (first word after the accumulators) = logging enabled
    const unsigned long long i = atomicAdd(acc + DBG_LOG_N, 1ull);
    if (i >= DBG_LOG_CAP) return;
    unsigned long long* e = acc + DBG_SLOTS + 8 + i * 4;
    e[0] = sm_id(); e[1] = (unsigned long long)kernel_id; e[2] = t0; e[3] = global_ns();
}

struct Bars {
    uint64_t full[RING], empty[RING];
    uint64_t a_full[2];   // row warps -> MMA: column pass p of the operand image written (count 512)
    uint64_t s_full[2];   // MMA -> row warps: accumulator S0 / S1 complete (tcgen05.commit)
    uint64_t a_free[2];   // MMA -> row warps (k_conv): the MMAs reading column pass p of the image have completed
    uint64_t x_full;      // producer -> row warps (k_enc): the tile's residual stream has landed in the image area
    uint64_t s_free;      // row warps -> MMA (k_enc): accumulator T has been read by every row thread (count 512)
    uint32_t tmem_base;
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------------------------
// fp32 -> (hi, lo) fp16 split and the swizzled operand-image stores
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// columns [c0, c0+32) (c0 % 32 == 0) of row r into a (hi, lo) pair of operand images made of 64-column slabs
template <uint32_t SLABB = SLAB_BYTES>
__device__ __forceinline__ void store_row32_split(uint8_t* img_hi, uint8_t* img_lo, uint32_t r, uint32_t c0,
                                                  const float (&v)[32]) {
    const uint32_t slab = (c0 >> 6) * SLABB;
    const uint32_t j0 = (c0 & 63) >> 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 h, l;
        split8(&v[8 * j], h, l);
        const uint32_t off = slab + slab_chunk_off(r, j0 + j);
        *reinterpret_cast<uint4*>(img_hi + off) = h;
        *reinterpret_cast<uint4*>(img_lo + off) = l;
    }
}

// the same, hi part only (operands the precision map keeps as ONE fp16 value)
__device__ __forceinline__ void store_row32_hi(uint8_t* img_hi, uint32_t r, uint32_t c0, const float (&v)[32]) {
    const uint32_t slab = (c0 >> 6) * SLAB_BYTES;
    const uint32_t j0 = (c0 & 63) >> 3;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(img_hi + slab + slab_chunk_off(r, j0 + j)) = pack8_f16(&v[8 * j]);
}

// columns [c0, c0+16) (c0 % 16 == 0)
__device__ __forceinline__ void store_row16_split(uint8_t* img_hi, uint8_t* img_lo, uint32_t r, uint32_t c0,
                                                  const float (&v)[16]) {
    const uint32_t slab = (c0 >> 6) * SLAB_BYTES;
    const uint32_t j0 = (c0 & 63) >> 3;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        uint4 h, l;
        split8(&v[8 * j], h, l);
        const uint32_t off = slab + slab_chunk_off(r, j0 + j);
        *reinterpret_cast<uint4*>(img_hi + off) = h;
        *reinterpret_cast<uint4*>(img_lo + off) = l;
    }
}

// ---------------------------------------------------------------------------------------------------------
// weight image stage order
// ---------------------------------------------------------------------------------------------------------
// stage order of one 256x256 block W[n][k]: for ks (k-slab of 64): hi n<128 | hi n>=128 | lo n<128 | lo n>=128
__host__ __device__ __forceinline__ size_t gemm_stage_off(int ks, int lo, int nh) {
    return (size_t)((ks * 2 + lo) * 2 + nh) * STAGE_HALFS;
}

// ---------------------------------------------------------------------------------------------------------
// tile bookkeeping
// ---------------------------------------------------------------------------------------------------------
struct TileGeom {
    int B, L1, L2, T1, T2;          // tiles per image: T = ceil(L/128)
    __host__ __device__ int tiles() const { return B * (T1 + T2); }
};
struct TileInfo { int set, b, ti, L, T, img, valid, first_tile_of_img; };
__host__ __device__ __forceinline__ TileInfo tile_info(const TileGeom& g, int t) {
    TileInfo ti;
    if (t < g.B * g.T1) { ti.set = 0; ti.b = t / g.T1; ti.ti = t % g.T1; ti.L = g.L1; ti.T = g.T1; ti.first_tile_of_img = ti.b * g.T1; }
    else { const int u = t - g.B * g.T1; ti.set = 1; ti.b = u / g.T2; ti.ti = u % g.T2; ti.L = g.L2; ti.T = g.T2;
           ti.first_tile_of_img = g.B * g.T1 + ti.b * g.T2; }
    ti.img = ti.set * g.B + ti.b;
    ti.valid = ti.L - ti.ti * TILE < TILE ? ti.L - ti.ti * TILE : TILE;
    return ti;
}
// Row mapping of the ENCODER tiles (k_enc, k_fold, k_sum_partials).  flat == 0: the per-image tiles of TileGeom.
// flat == 1: the images of a set are concatenated, each padded to Lp = round_up(L, 16) rows, and the B*Lp rows are cut
// into 128-row tiles: no per-image padding to a multiple of 128 (400 tokens: 3.125 tiles instead of 4).  With
// Lp >= 128 a tile holds rows of at most two consecutive images and the boundary is a multiple of 16 rows (one MMA
// k-step of the K^T V product).  The head kernels keep per-image tiles (k_retile converts the encoder output).
struct EncGeom {
    int flat, Lp1, Lp2, F1, F2;     // F: flat tiles per set
};
struct EncTile { int set, L, Lp, B, b0, l0, split, two; };
// rows [0, split) of the tile belong to image b0 (tokens l0 ..), rows [split, 128) to image b0 + 1 (tokens 0 ..)
__host__ __device__ __forceinline__ EncTile enc_tile(const TileGeom& g, const EncGeom& eg, int t) {
    EncTile e;
    e.B = g.B;
    if (!eg.flat) {
        const TileInfo ti = tile_info(g, t);
        e.set = ti.set; e.L = ti.L; e.Lp = ti.L; e.b0 = ti.b; e.l0 = ti.ti * 128; e.split = 128; e.two = 0;
    } else {
        e.set = t >= eg.F1 ? 1 : 0;
        const int u = t - e.set * eg.F1;
        e.L = e.set ? g.L2 : g.L1; e.Lp = e.set ? eg.Lp2 : eg.Lp1;
        const int base = u * 128;
        e.b0 = base / e.Lp; e.l0 = base - e.b0 * e.Lp;
        e.split = e.Lp - e.l0 < 128 ? e.Lp - e.l0 : 128;
        e.two = (e.split < 128 && e.b0 + 1 < g.B) ? 1 : 0;
    }
    return e;
}
// the partial summaries of image (set, b): n = enc_parts(...), then enc_part_index(..., i) for i < n, in a fixed order.
// ppt = partial slots per tile (flat: 2 = one per image of a tile)
__host__ __device__ __forceinline__ int enc_parts(const TileGeom& g, const EncGeom& eg, int ppt, int set, int b) {
    if (!eg.flat) return (set == 0 ? g.T1 : g.T2) * ppt;
    const int L = set ? g.L2 : g.L1, Lp = set ? eg.Lp2 : eg.Lp1;
    return (b * Lp + L - 1) / 128 - (b * Lp) / 128 + 1;
}
__host__ __device__ __forceinline__ int enc_part_index(const TileGeom& g, const EncGeom& eg, int ppt, int set, int b, int i) {
    if (!eg.flat) return (set == 0 ? b * g.T1 : g.B * g.T1 + b * g.T2) * ppt + i;
    const int Lp = set ? eg.Lp2 : eg.Lp1;
    const int u = (b * Lp) / 128 + i;                       // flat tile inside the set
    const int slot = b - (u * 128) / Lp;                    // 0: the tile starts inside this image, 1: inside the previous one
    return ((set ? eg.F1 : 0) + u) * 2 + slot;
}
// tile-blocked fp32 layout [tile][64 col-quads][128 rows][4]: a warp's rows read/write one col-quad coalesced
__host__ __device__ __forceinline__ size_t xt_off(int tile, int quad, int r) { return (((size_t)tile * 64 + quad) * TILE + r) * 4; }

// single-instruction special functions (flush-to-zero forms: no denormal fix-up code around the MUFU)
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// elu(x) + 1 (linear_attention.py:12-13)
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x + 1.f : ex2_ftz(x * 1.4426950408889634f); }
// nn.GELU (erf form, transformer.py:93).  erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7), branch-free:
// ~13 instructions instead of erff's divergent ~35; the result error (<= 3e-7 |x|) is far inside the parity budget.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = rcp_ftz(fmaf(0.3275911f, z, 1.f));
    float poly = fmaf(t, 1.061405429f, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    const float erf_abs = fmaf(-poly, ex2_ftz(z * z * -1.4426950408889634f), 1.f);
    const float hx = 0.5f * x;
    return fmaf(hx, copysignf(erf_abs, x), hx);
}

// ---------------------------------------------------------------------------------------------------------
// device building blocks shared by k_enc and k_conv
// ---------------------------------------------------------------------------------------------------------
// barriers + TMEM allocation (whole CTA); returns the TMEM base address
__device__ __forceinline__ uint32_t cta_setup(Bars* bars, int alloc_warp) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->a_full[i], N_ROW_THREADS);
            mbar_init(&bars->s_full[i], 1);
            mbar_init(&bars->a_free[i], 1);
        }
        mbar_init(&bars->x_full, 1);
        mbar_init(&bars->s_free, N_ROW_THREADS);
        fence_mbar_init();
    }
    if ((int)(threadIdx.x >> 5) == alloc_warp) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return bars->tmem_base;
}
// producer: `nunits` 32 KB units (two adjacent 16 KB stages = one [256 N x 64 K] B tile) global -> ring, ONE bulk copy
// per unit.  The issuing thread pays ~480 cycles per copy whatever its size (measured, tools/bulk_bw.cu: 16 KB copies
// stream at 34 B/cycle/SM, 32 KB copies at 68), and the 3-term MMAs consume 43 B/cycle: 16 KB copies starve the
// tensor core.  unit_stride: distance between consecutive units in stages (2 = every unit of a GEMM image: hi and lo;
// 4 = the hi units only, for GEMMs whose weight operand is not split).  Only the even stage's full barrier is used;
// both empty barriers are still committed by the consumer.
__device__ __forceinline__ void ring_stream(uint8_t* smem, Bars* bars, int* flag, uint32_t& g, const __half* src, int nunits,
                                            int unit_stride = 2) {
    for (int i = 0; i < nunits; ++i, g += 2) {
        const int st = g % RING;
        mbar_wait(&bars->empty[st], ((g / RING) & 1) ^ 1, flag);
        mbar_wait(&bars->empty[st + 1], ((g / RING) & 1) ^ 1, flag);
        mbar_arrive_expect_tx(&bars->full[st], 2 * STAGE_BYTES);
        bulk_g2s(smem + SM_RING + st * STAGE_BYTES, src + (size_t)i * unit_stride * STAGE_HALFS, 2 * STAGE_BYTES, &bars->full[st]);
    }
}
struct MmaState { uint32_t g = 0, na0 = 0, na1 = 0; long long t_a = 0, t_ring = 0; };   // t_*: cycles spent waiting (profiling aid)
__device__ __forceinline__ void mma_wait_a(Bars* bars, int* flag, MmaState& ms, int pass) {
    const long long t0 = clock64();
    mbar_wait(&bars->a_full[pass], (pass ? ms.na1++ : ms.na0++) & 1, flag);
    ms.t_a += clock64() - t0;
    tc_fence_after();
}
// Terms of a split product a.w = (a_hi + a_lo).(w_hi + w_lo): which ones a GEMM issues is decided per contraction by
// the precision map (tests/precision_map.py); a_lo.w_lo (2^-22 relative) is never issued.
enum { T_HH = 1, T_LH = 2, T_HL = 4, T_ALL = 7 };     // a_hi.w_hi | a_lo.w_hi | a_hi.w_lo
// D[128 x 256] (tmem columns d..d+255) (+)= A[128 x 256] . W^T; consumes 4 ring units (w_hi), or 8 with T_HL (w_hi,
// w_lo alternating).  wait: the operand image is (re)written for this GEMM -> wait for column pass 0 before k-slab 0
// and pass 1 before k-slab 2.  signal_free: commit a_free[p] once the MMAs reading pass p have been issued (k_conv).
__device__ __forceinline__ void gemm_issue(uint32_t smem_base, Bars* bars, int* flag, MmaState& ms, uint32_t d,
                                           bool accumulate, bool wait, bool signal_free, int terms = T_ALL) {
    for (int ks = 0; ks < 4; ++ks) {
        if (wait && ks == 0) mma_wait_a(bars, flag, ms, 0);
        if (wait && ks == 2) mma_wait_a(bars, flag, ms, 1);
        const uint32_t a_hi = smem_base + SM_AHI + ks * SLAB_BYTES;
        const uint32_t a_lo = smem_base + SM_ALO + ks * SLAB_BYTES;
        {   // w_hi: two adjacent stages form the [256 x 64] B tile
            const int st = ms.g % RING;
            const long long t0 = clock64();
            mbar_wait(&bars->full[st], (ms.g / RING) & 1, flag);
            ms.t_ring += clock64() - t0;
            tc_fence_after();
            const uint32_t b = smem_base + SM_RING + st * STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_f16(d, umma_desc(a_hi + k * 32, 16, ATOM_BYTES), umma_desc(b + k * 32, 16, ATOM_BYTES),
                         IDESC_N256, (accumulate || ks > 0 || k > 0) ? 1u : 0u);
            if (terms & T_LH) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(d, umma_desc(a_lo + k * 32, 16, ATOM_BYTES), umma_desc(b + k * 32, 16, ATOM_BYTES),
                             IDESC_N256, 1u);
            }
            umma_commit(&bars->empty[st]);
            umma_commit(&bars->empty[st + 1]);
            ms.g += 2;
        }
        if (terms & T_HL) {   // w_lo
            const int st = ms.g % RING;
            const long long t0 = clock64();
            mbar_wait(&bars->full[st], (ms.g / RING) & 1, flag);
            ms.t_ring += clock64() - t0;
            tc_fence_after();
            const uint32_t b = smem_base + SM_RING + st * STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_f16(d, umma_desc(a_hi + k * 32, 16, ATOM_BYTES), umma_desc(b + k * 32, 16, ATOM_BYTES),
                         IDESC_N256, 1u);
            umma_commit(&bars->empty[st]);
            umma_commit(&bars->empty[st + 1]);
            ms.g += 2;
        }
        if (signal_free && (ks & 1)) umma_commit(&bars->a_free[ks >> 1]);
    }
}

}  // namespace oetr
