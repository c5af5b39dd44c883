// Full (softmax) attention on the tensor cores: attention_mode = 'full' of the reference's QueryTransformer
// (src/models/transformer.py:288-296 -> FullAttention, src/models/linear_attention.py:59-87) with OETR_PREC_FP16.
// Part of the tcgen05 path; compiled into tc_kernels.cu.
//
// Nothing can be folded here (softmax sits between the two contractions), so a layer is two kernels:
//   k_proj_mlp : per 128-token tile of ONE image (per-image tiles): [merge + MLP of layer i from the attention output
//                image O: x += O Wm^T ; x += W2 gelu(W1 LN2(x))]  then  [projections of layer i+1: q = (LNq(x)+pos) Wq^T
//                scaled by log2(e)/sqrt(32), k = (LNkv(x)+pos) Wk^T, v = ... Wv^T], written as fp16 (hi, lo) operand
//                images in the 128-byte-swizzled slab layout the attention kernel streams.  Same machinery as k_enc
//                (residual stream in TMEM, tensor-core residual adds); every product is a 3-term split (the softmax
//                amplifies operand rounding: tests/precision_map_full.py).
//   k_attn     : per (query tile, pair of heads): S = Q_h K_h^T on the tensor cores (M = 128 queries, N = 64 keys per
//                chunk, K = 32), row softmax by one thread per query row straight out of TMEM (two passes over the
//                key chunks: row max, then p = exp2(s - max) and the row sum, so no accumulator is ever rescaled),
//                O_h += P V on the tensor cores (P written to shared memory as the A operand, V chunks MN-major),
//                O / rowsum written as the (hi, lo) A-operand image of the merge GEMM.  Cross layers read the partner
//                image's k, v (all of them: SURVEY.md Appendix A.2).
#pragma once
#include "tc_tiles.cuh"

namespace oetr {
using namespace tc;

constexpr float QK_SCALE_LOG2 = 0.17677669529663687f * 1.4426950408889634f;      // 1/sqrt(32) * log2(e)
// operand images exchanged through global memory: per tile [hi | lo][4 slabs][128 rows x 64 cols] fp16 = 128 KB
constexpr size_t TILE_IMG_HALFS = 2 * (size_t)IMG_BYTES / 2;

// ---------------------------------------------------------------------------------------------------------
// k_proj_mlp
// ---------------------------------------------------------------------------------------------------------
struct ProjParams {
    TileGeom g;
    const float *feat1, *feat2;     // NCHW inputs, read when load_feat
    float* xt;                      // tile-blocked residual stream (per-image tiles)
    const float *post1, *post2;
    int load_feat, do_merge, do_proj;
    const __half* oimg;             // [tiles][TILE_IMG_HALFS] attention output (A operand of the merge GEMM)
    const __half* w_merge;          // Wm GEMM image
    const __half* w_mlp;            // W1a | W1b | W2a | W2b
    const __half* w_q;              // Wq
    const __half* w_kv;             // Wv | Wk
    __half *qimg, *kimg, *vimg;     // [tiles][TILE_IMG_HALFS] projections of the next layer
    int* flag;
    float ln2_g[C], ln2_b[C], lnq_g[C], lnq_b[C], lnkv_g[C], lnkv_b[C];
};

__global__ void __launch_bounds__(N_THREADS, 1) k_proj_mlp(const __grid_constant__ ProjParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const TileInfo ti = tile_info(p.g, blockIdx.x);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t tmem = cta_setup(bars, WARP_PRODUCER);
    const uint32_t T = tmem, XA = tmem + 256;

    if (warp == WARP_PRODUCER) {
        if (lane == 0) {
            if (p.do_merge) {     // attention output image of this tile -> operand image area (hi | lo contiguous)
                mbar_arrive_expect_tx(&bars->x_full, 2 * IMG_BYTES);
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.oimg + (size_t)blockIdx.x * TILE_IMG_HALFS);
#pragma unroll 1
                for (int i = 0; i < 4; ++i)
                    bulk_g2s(smem + SM_AHI + i * (IMG_BYTES / 2), src + (size_t)i * (IMG_BYTES / 2), IMG_BYTES / 2, &bars->x_full);
            }
            uint32_t g = 0;
            if (p.do_merge) {
                ring_stream(smem, bars, p.flag, g, p.w_merge, 8, 2);
                ring_stream(smem, bars, p.flag, g, p.w_mlp, 32, 2);
            }
            if (p.do_proj) {
                ring_stream(smem, bars, p.flag, g, p.w_q, 8, 2);
                ring_stream(smem, bars, p.flag, g, p.w_kv, 16, 2);
            }
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        if (lane == 0) {
            MmaState ms;
            auto gemm = [&](uint32_t d, bool accumulate, bool wait) { gemm_issue(smem_base, bars, p.flag, ms, d, accumulate, wait, false, T_ALL); };
            if (p.do_merge) {
                mbar_wait(&bars->x_full, 0, p.flag);          // O image landed; a_full: the residual stream is in XA
                tc_fence_after();
                gemm(XA, true, true);   umma_commit(&bars->s_full[1]);     // x += O Wm^T
                gemm(T, false, true);   umma_commit(&bars->s_full[0]);     // h_a = LN2(x) W1a^T
                mbar_wait(&bars->s_free, 0, p.flag);
                tc_fence_after();
                gemm(T, false, false);  umma_commit(&bars->s_full[0]);     // h_b
                gemm(XA, true, true);   umma_commit(&bars->s_full[1]);     // x += gelu(h_a) W2a^T
                gemm(XA, true, true);   umma_commit(&bars->s_full[1]);     // x += gelu(h_b) W2b^T
            }
            if (p.do_proj) {
                gemm(T, false, true);   umma_commit(&bars->s_full[0]);     // q
                gemm(T, false, true);   umma_commit(&bars->s_full[0]);     // v (new image: LNkv(x)+pos)
                gemm(XA, false, false); umma_commit(&bars->s_full[1]);     // k
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3, cq = warp >> 2;
        const int r = q * 32 + lane;
        const bool valid = r < ti.valid;
        const int pl = valid ? ti.ti * TILE + r : 0;           // token inside the image = row of the position table
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint8_t* img_hi = smem + SM_AHI;
        uint8_t* img_lo = smem + SM_ALO;
        const float* post = (ti.set == 0 ? p.post1 : p.post2);
        uint32_t ns0 = 0, ns1 = 0;
        auto wait_s = [&](int b) {
            mbar_wait(&bars->s_full[b], (b ? ns1++ : ns0++) & 1, p.flag);
            tc_fence_after();
        };
        auto publish = [&](int pass) {
            tc_fence_before();
            fence_async_smem();
            mbar_arrive(&bars->a_full[pass]);
        };
        auto load_acc = [&](uint32_t acc, float (&v)[2][32]) { tmem_ld32x2(acc + lane_addr + cq * 32, v[0], v[1]); };
        const bool pos_chunk = cq < 2;                     // see k_enc: the table only matters for channels < 64
        auto load_pos = [&](float4 (&ps)[8]) {
            if (pos_chunk) {
#pragma unroll
                for (int jq = 0; jq < 8; ++jq)
                    ps[jq] = __ldg(reinterpret_cast<const float4*>(post + xt_off(pl >> 7, cq * 8 + jq, pl & 127)));
            } else {
#pragma unroll
                for (int jq = 0; jq < 8; ++jq) ps[jq] = make_float4(0.f, 1.f, 0.f, 1.f);
            }
        };
        auto row_stats = [&](const float (&x)[2][32], float& mean, float& rstd) {      // see k_enc
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) s += x[0][e] + x[1][e];
            const float m_i = s * (1.f / 64.f);
            float m2 = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const float d0 = x[0][e] - m_i, d1 = x[1][e] - m_i;
                m2 = fmaf(d0, d0, m2);
                m2 = fmaf(d1, d1, m2);
            }
            tmem_st2(T + lane_addr + cq * 32, m_i, m2);
            tmem_st_wait();
            tc_fence_before();
            named_bar_sync(2 + q, 128);
            tc_fence_after();
            float v[8];
            tmem_ld2x4(T + lane_addr, v);
            mean = (v[0] + v[2] + v[4] + v[6]) * 0.25f;
            float M2 = (v[1] + v[3]) + (v[5] + v[7]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float d = v[2 * i] - mean;
                M2 = fmaf(64.f * d, d, M2);
            }
            rstd = rsqrtf(M2 * (1.f / C) + LN_EPS);
        };
        auto ln_image = [&](float (&x)[2][32], const float* __restrict__ gamma, const float* __restrict__ beta, bool with_pos) {
            float4 ps[8];
            if (with_pos) load_pos(ps);
            float mean, rstd;
            row_stats(x, mean, rstd);
            const float shift = -mean * rstd;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;
#pragma unroll
                for (int jq = 0; jq < 8; ++jq) {
                    const float4 pz = with_pos ? ps[jq] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const int c = c0 + jq * 4;
                    x[pass][jq * 4 + 0] = fmaf(fmaf(x[pass][jq * 4 + 0], rstd, shift), gamma[c + 0], beta[c + 0] + pz.x);
                    x[pass][jq * 4 + 1] = fmaf(fmaf(x[pass][jq * 4 + 1], rstd, shift), gamma[c + 1], beta[c + 1] + pz.y);
                    x[pass][jq * 4 + 2] = fmaf(fmaf(x[pass][jq * 4 + 2], rstd, shift), gamma[c + 2], beta[c + 2] + pz.z);
                    x[pass][jq * 4 + 3] = fmaf(fmaf(x[pass][jq * 4 + 3], rstd, shift), gamma[c + 3], beta[c + 3] + pz.w);
                }
                if (with_pos && pass == 0) {
#pragma unroll
                    for (int jq = 0; jq < 8; ++jq) ps[jq] = make_float4(0.f, 1.f, 0.f, 1.f);
                }
                store_row32_split(img_hi, img_lo, r, c0, x[pass]);
                publish(pass);
            }
        };
        auto load_x = [&](float (&x)[2][32]) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;
#pragma unroll
                for (int jq = 0; jq < 8; ++jq) {
                    const float4 v = *reinterpret_cast<const float4*>(p.xt + xt_off(blockIdx.x, (c0 >> 2) + jq, r));
                    x[pass][jq * 4 + 0] = v.x; x[pass][jq * 4 + 1] = v.y; x[pass][jq * 4 + 2] = v.z; x[pass][jq * 4 + 3] = v.w;
                }
            }
        };
        auto store_x = [&](const float (&x)[2][32]) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;
#pragma unroll
                for (int jq = 0; jq < 8; ++jq)
                    *reinterpret_cast<float4*>(p.xt + xt_off(blockIdx.x, (c0 >> 2) + jq, r)) =
                        make_float4(x[pass][jq * 4], x[pass][jq * 4 + 1], x[pass][jq * 4 + 2], x[pass][jq * 4 + 3]);
            }
        };
        // (hi, lo) operand image of this tile in global memory, same swizzled slab layout as the shared-memory image
        auto store_global_image = [&](__half* base, const float (&v)[2][32], float scale, bool zero) {
            uint8_t* hi = reinterpret_cast<uint8_t*>(base + (size_t)blockIdx.x * TILE_IMG_HALFS);
            uint8_t* lo = hi + IMG_BYTES;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                float w[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) w[e] = zero ? 0.f : v[pass][e] * scale;
                store_row32_split(hi, lo, r, pass * 128 + cq * 32, w);
            }
        };

        float x[2][32];
        if (p.load_feat) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;
                const float* feat = ti.set == 0 ? p.feat1 : p.feat2;
                const float* f = feat + ((size_t)(valid ? ti.b : 0) * C + c0) * ti.L + pl;
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] = valid ? f[(size_t)e * ti.L] : 0.f;
            }
        } else {
            load_x(x);
        }
        if (p.do_merge) {
            tmem_st32(XA + lane_addr + cq * 32, x[0]);
            tmem_st32(XA + lane_addr + 128 + cq * 32, x[1]);
            tmem_st_wait();
            publish(0);                                    // "the residual stream is in XA": the merge GEMM may accumulate
            publish(1);
            wait_s(1);
            load_acc(XA, x);                               // x + O Wm^T
            ln_image(x, p.ln2_g, p.ln2_b, false);
            wait_s(0);
            load_acc(T, x);
            tc_fence_before();
            mbar_arrive(&bars->s_free);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass)
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] = gelu_erf(x[pass][e]);
            wait_s(0);
            store_row32_split(img_hi, img_lo, r, cq * 32, x[0]);
            store_row32_split(img_hi, img_lo, r, 128 + cq * 32, x[1]);
            load_acc(T, x);
            publish(0);
            publish(1);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass)
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] = gelu_erf(x[pass][e]);
            wait_s(1);
            store_row32_split(img_hi, img_lo, r, cq * 32, x[0]);
            publish(0);
            store_row32_split(img_hi, img_lo, r, 128 + cq * 32, x[1]);
            publish(1);
            wait_s(1);
            load_acc(XA, x);
        }
        if (p.do_merge || p.load_feat) store_x(x);
        if (p.do_proj) {
            ln_image(x, p.lnq_g, p.lnq_b, true);
            wait_s(0);
            load_acc(T, x);
            store_global_image(p.qimg, x, QK_SCALE_LOG2, false);
            load_x(x);                                     // this thread's own stores (or an earlier launch's)
            ln_image(x, p.lnkv_g, p.lnkv_b, true);
            wait_s(0);
            load_acc(T, x);
            store_global_image(p.vimg, x, 1.f, !valid);    // rows beyond the image: zeros (their softmax weight is 0)
            wait_s(1);
            load_acc(XA, x);
            store_global_image(p.kimg, x, 1.f, !valid);
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_PRODUCER) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------
// k_attn
// ---------------------------------------------------------------------------------------------------------
constexpr int AT_THREADS = 192;                 // warps 0-3: one thread per query row; warp 4: loads; warp 5: MMA issue
constexpr int AT_CHUNK = 64;                    // keys per chunk
constexpr int AT_RING = 5;                      // stages of one (hi, lo) chunk of K or V: 2 x 8 KB
constexpr uint32_t AT_HALF = AT_CHUNK * 128;    // one [64 rows x 64 cols] fp16 half slab: 8 KB
constexpr uint32_t AT_Q = 0;                    // Q slab pair (hi 16 KB | lo 16 KB)
constexpr uint32_t AT_P = AT_Q + 2 * SLAB_BYTES;     // P slab pair
constexpr uint32_t AT_RING_OFF = AT_P + 2 * SLAB_BYTES;
constexpr uint32_t AT_BAR = AT_RING_OFF + AT_RING * 2 * AT_HALF;
constexpr uint32_t AT_TOTAL = AT_BAR + 256;
constexpr uint32_t IDESC_QK = umma_idesc_f16(128, 64, 0, 0);        // S[128 x 64] = Q[128 x 16] . K[64 x 16]^T
constexpr uint32_t IDESC_PV = umma_idesc_f16(128, 64, 0, 1);        // O[128 x 64] += P[128 x 16] . V[16 x 64]   (V MN-major)

struct AttnBars {
    uint64_t full[AT_RING], empty[AT_RING];
    uint64_t q_full;
    uint64_t s_full[2], s_free[2];      // S buffer b: MMA -> rows (commit) / rows -> MMA (128 arrivals)
    uint64_t p_full, p_free;            // P slab: rows -> MMA (128 arrivals) / MMA -> rows (commit)
    uint64_t o_full, o_free;            // O accumulator of a head: MMA -> rows (commit) / rows -> MMA (128 arrivals)
    uint32_t tmem_base, pad;
};
static_assert(sizeof(AttnBars) <= 256, "AttnBars must fit its reservation");

struct AttnParams {
    TileGeom g;
    const __half *qimg, *kimg, *vimg;   // [tiles][TILE_IMG_HALFS]
    __half* oimg;                       // [tiles][TILE_IMG_HALFS]
    int cross;                          // keys / values of the partner image (transformer.py:354-358)
    int* flag;
};

__global__ void __launch_bounds__(AT_THREADS, 1) k_attn(const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    AttnBars* bars = reinterpret_cast<AttnBars*>(smem + AT_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, pair = blockIdx.y;                     // head pair: heads 2*pair, 2*pair + 1 = slab `pair`
    const TileInfo ti = tile_info(p.g, tile);
    // the source image: its tiles and its length
    const int s_set = p.cross ? 1 - ti.set : ti.set;
    const int S = s_set == 0 ? p.g.L1 : p.g.L2;
    const int s_first = s_set == 0 ? ti.b * p.g.T1 : p.g.B * p.g.T1 + ti.b * p.g.T2;
    const int nchunks = (S + AT_CHUNK - 1) / AT_CHUNK;
    const uint32_t smem_base = smem_u32(smem);

    if (tid == 0) {
        for (int i = 0; i < AT_RING; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        mbar_init(&bars->q_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_free[i], 128); }
        mbar_init(&bars->p_full, 128); mbar_init(&bars->p_free, 1);
        mbar_init(&bars->o_full, 1);   mbar_init(&bars->o_free, 128);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc(&bars->tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t SB[2] = {tmem, tmem + 64};
    const uint32_t OA = tmem + 128;

    // chunk c of the source image, slab `pair`: rows [64*(c&1), +64) of source tile (c >> 1); hi at +0, lo at +IMG_BYTES
    auto chunk_src = [&](const __half* img, int c) {
        return reinterpret_cast<const uint8_t*>(img + (size_t)(s_first + (c >> 1)) * TILE_IMG_HALFS) + pair * SLAB_BYTES + (c & 1) * AT_HALF;
    };

    if (warp == 4) {
        // ------------------------------------------------------------------ loads: Q once, then K / V chunks in MMA order
        if (lane == 0) {
            const uint8_t* qsrc = reinterpret_cast<const uint8_t*>(p.qimg + (size_t)tile * TILE_IMG_HALFS) + pair * SLAB_BYTES;
            mbar_arrive_expect_tx(&bars->q_full, 2 * SLAB_BYTES);
            bulk_g2s(smem + AT_Q, qsrc, SLAB_BYTES, &bars->q_full);
            bulk_g2s(smem + AT_Q + SLAB_BYTES, qsrc + IMG_BYTES, SLAB_BYTES, &bars->q_full);
            uint32_t g = 0;
            auto push = [&](const __half* img, int c) {
                const int st = g % AT_RING;
                mbar_wait(&bars->empty[st], ((g / AT_RING) & 1) ^ 1, p.flag);
                mbar_arrive_expect_tx(&bars->full[st], 2 * AT_HALF);
                const uint8_t* src = chunk_src(img, c);
                bulk_g2s(smem + AT_RING_OFF + st * 2 * AT_HALF, src, AT_HALF, &bars->full[st]);
                bulk_g2s(smem + AT_RING_OFF + st * 2 * AT_HALF + AT_HALF, src + IMG_BYTES, AT_HALF, &bars->full[st]);
                ++g;
            };
            for (int hh = 0; hh < 2; ++hh) {
                for (int c = 0; c < nchunks; ++c) push(p.kimg, c);                       // pass 1: row maxima
                // pass 2 in the MMA lane's consumption order: K0, (K1, V0), (K2, V1), ..., V_last
                push(p.kimg, 0);
                for (int c = 0; c < nchunks; ++c) {
                    if (c + 1 < nchunks) push(p.kimg, c + 1);
                    push(p.vimg, c);
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issue
        if (lane == 0) {
            uint32_t g = 0, nsf[2] = {0, 0}, npf = 0, nof = 0;
            uint32_t sb = 0;                                            // S buffer of the next QK product
            mbar_wait(&bars->q_full, 0, p.flag);
            tc_fence_after();
            auto ring_wait = [&]() {
                const int st = g % AT_RING;
                mbar_wait(&bars->full[st], (g / AT_RING) & 1, p.flag);
                tc_fence_after();
                return smem_base + AT_RING_OFF + st * 2 * AT_HALF;
            };
            auto ring_release = [&]() { umma_commit(&bars->empty[g % AT_RING]); ++g; };
            // S[sb] = Q_h K_c^T, 3-term split: q_hi k_hi + q_lo k_hi + q_hi k_lo ; head hh = columns [32*hh, +32) of the slab
            auto qk = [&](int hh, bool first_use) {
                const uint32_t kb = ring_wait();
                if (!first_use) {                                       // the rows have read the previous contents of S[sb]
                    mbar_wait(&bars->s_free[sb], nsf[sb]++ & 1, p.flag);
                    tc_fence_after();
                }
                const uint32_t q_hi = smem_base + AT_Q + hh * 64, q_lo = q_hi + SLAB_BYTES;
                const uint32_t k_hi = kb + hh * 64, k_lo = k_hi + AT_HALF;
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const uint32_t a = term == 1 ? q_lo : q_hi, b = term == 2 ? k_lo : k_hi;
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        umma_f16(SB[sb], umma_desc(a + k * 32, 16, ATOM_BYTES), umma_desc(b + k * 32, 16, ATOM_BYTES), IDESC_QK,
                                 (term | k) ? 1u : 0u);
                }
                umma_commit(&bars->s_full[sb]);
                ring_release();
                sb ^= 1;
            };
            uint32_t s_uses = 0;
            for (int hh = 0; hh < 2; ++hh) {
                for (int c = 0; c < nchunks; ++c) { qk(hh, s_uses < 2); ++s_uses; }       // pass 1
                qk(hh, s_uses < 2); ++s_uses;                                             // pass 2, chunk 0
                for (int c = 0; c < nchunks; ++c) {
                    if (c + 1 < nchunks) { qk(hh, s_uses < 2); ++s_uses; }                // next chunk's scores under this chunk's softmax
                    // O += P_c V_c : P [128 x 64 keys] (hi, lo), V chunk [64 keys x 64 dims] MN-major (hi, lo)
                    const uint32_t vb = ring_wait();
                    mbar_wait(&bars->p_full, npf++ & 1, p.flag);
                    tc_fence_after();
                    if (c == 0 && hh > 0) {                             // the previous head's O has been read
                        mbar_wait(&bars->o_free, nof++ & 1, p.flag);
                        tc_fence_after();
                    }
                    const uint32_t p_hi = smem_base + AT_P, p_lo = p_hi + SLAB_BYTES;
                    const uint32_t v_hi = vb, v_lo = vb + AT_HALF;
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t a = term == 1 ? p_lo : p_hi, b = term == 2 ? v_lo : v_hi;
#pragma unroll
                        for (int k = 0; k < AT_CHUNK / 16; ++k)
                            umma_f16(OA, umma_desc(a + k * 32, 16, ATOM_BYTES), umma_desc(b + k * 2048, SLAB_BYTES, ATOM_BYTES), IDESC_PV,
                                     (c | term | k) ? 1u : 0u);
                    }
                    umma_commit(&bars->p_free);
                    ring_release();
                }
                umma_commit(&bars->o_full);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ softmax rows: thread <-> query row
        const int r = warp * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        uint32_t nsu[2] = {0, 0}, npf = 0, nou = 0;
        uint32_t sb = 0;
        uint8_t* p_hi = smem + AT_P;
        uint8_t* p_lo = p_hi + SLAB_BYTES;
        uint8_t* o_hi = reinterpret_cast<uint8_t*>(p.oimg + (size_t)tile * TILE_IMG_HALFS) + pair * SLAB_BYTES;
        uint8_t* o_lo = o_hi + IMG_BYTES;
        for (int hh = 0; hh < 2; ++hh) {
            // pass 1: row maximum of the (log2-domain) scores over the valid keys
            float m = -INFINITY;
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(&bars->s_full[sb], nsu[sb]++ & 1, p.flag);
                tc_fence_after();
                float s0[32], s1[32];
                tmem_ld32x2_adj(SB[sb] + lane_addr, s0, s1);
                tc_fence_before();
                mbar_arrive(&bars->s_free[sb]);
                const int nvalid = S - c * AT_CHUNK;                    // keys of this chunk that exist
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    if (e < nvalid) m = fmaxf(m, s0[e]);
                    if (e + 32 < nvalid) m = fmaxf(m, s1[e]);
                }
                sb ^= 1;
            }
            // pass 2: p = exp2(s - m), row sum in fp32 from the unrounded p, P operand image, O += P V on the tensor cores
            float l = 0.f;
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(&bars->s_full[sb], nsu[sb]++ & 1, p.flag);
                tc_fence_after();
                float s0[32], s1[32];
                tmem_ld32x2_adj(SB[sb] + lane_addr, s0, s1);
                tc_fence_before();
                mbar_arrive(&bars->s_free[sb]);
                const int nvalid = S - c * AT_CHUNK;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    s0[e] = e < nvalid ? ex2_ftz(s0[e] - m) : 0.f;
                    s1[e] = e + 32 < nvalid ? ex2_ftz(s1[e] - m) : 0.f;
                    l += s0[e] + s1[e];
                }
                if (c > 0 || hh > 0) {                                  // the MMAs reading the previous P have completed
                    mbar_wait(&bars->p_free, npf++ & 1, p.flag);
                }
                store_row32_split(p_hi, p_lo, r, 0, s0);
                store_row32_split(p_hi, p_lo, r, 32, s1);
                fence_async_smem();
                mbar_arrive(&bars->p_full);
                sb ^= 1;
            }
            // this head's output: columns [32*hh, +32) of the pair's O accumulator, normalised, as the merge GEMM's operand image
            mbar_wait(&bars->o_full, nou++ & 1, p.flag);
            tc_fence_after();
            float o[32];
            tmem_ld32(OA + lane_addr + hh * 32, o);
            tc_fence_before();
            mbar_arrive(&bars->o_free);
            const float inv = 1.f / l;
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] *= inv;
            store_row32_split(o_hi, o_lo, r, hh * 32, o);
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 256);
}

}  // namespace oetr
