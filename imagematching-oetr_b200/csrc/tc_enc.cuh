// k_enc: one encoder layer (query phase of layer i + source phase of layer i+1) per 128-token tile, one CTA per tile.
// Part of the tcgen05 (OETR_PREC_FP16) path; compiled into tc_kernels.cu (one translation unit: kernels are
// launched from the host code there).
#pragma once
#include "tc_tiles.cuh"

namespace oetr {
using namespace tc;

// ---------------------------------------------------------------------------------------------------------
// k_enc
// ---------------------------------------------------------------------------------------------------------
// Operand precision per contraction (tests/precision_map.py; box error of the whole map 5e-5 of the image side):
//   q = (LNq(x)+pos) Wq^T            1 term   (a hi, w hi)          k = (LNkv(x)+pos) Wk^T   2 terms (a split, w hi)
//   v, merge, W1, W2, K^T V          3 terms  (both split)
//   decoder: v = x Wv^T              2 terms  (a hi, w split)       k = (x+pos) Wk^T 1 term; K^T V 1 term
struct EncParams {
    TileGeom g;
    EncGeom eg;                 // row mapping of the tiles
    const float* feat1;         // NCHW inputs, read when load_feat
    const float* feat2;
    float* xt;                  // tile-blocked residual stream (read unless load_feat; written when store_x)
    const float *post1, *post2; // tile-blocked positional rows of set 0 / set 1
    const float *mask1, *mask2; // nullable [B][L] float masks of set 0 / set 1: a row's mask scales its phi(q), phi(k)
                                // and v (linear_attention.py:36-41; k and v of a row always belong to its own image)
    int load_feat, store_x, do_q, do_kv;
    // query phase (encoder layer i)
    const __half* w_q;          // Wq                         (hi units only)
    const __half* w_mlp;        // W1a | W1b | W2a | W2b      (4 GEMM images)
    const __half* mimg;         // [2B images][GEMM_HALFS] folded merge weights of the source image
    const float* ksum;          // [2B images][256]
    int cross;                  // 1: the source is the partner image (transformer.py:354-358)
    // kv phase (encoder layer i+1, or a decoder layer's cross-attention when dec_mode)
    int dec_mode;               // decoder: k = (x+pos) Wk^T + bk, v = x Wv^T + bv, no LayerNorm (transformer.py:243-249)
    const __half* w_kv;         // Wv | Wk
    float* kv_part;             // [tiles][ppt][PART_FLOATS] per-tile partial summaries
    int* flag;
    unsigned long long* dbg_acc;   // nullable: global cycle accumulators (OETR_TIMING=1), see DBG_* in tc_tiles.cuh
    // L2 prefetch: every layer's weights are read once per forward, so without it each stage is a DRAM-latency
    // miss for the whole first wave.  The grid spreads these ranges (the NEXT launch's weights) in 16 KB pieces.
    const void* pf_ptr[3];
    uint32_t pf_bytes[3];
    // LayerNorm vectors and decoder biases travel in the kernel parameters (constant bank): every lane of a warp
    // reads the same column's value, and nothing has to be staged through shared memory (which is full)
    float lnq_g[C], lnq_b[C], ln2_g[C], ln2_b[C];
    float lnkv_g[C], lnkv_b[C];     // dec_mode: bv | bk
};

__global__ void __launch_bounds__(N_THREADS, 1) k_enc(const __grid_constant__ EncParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const EncTile et = enc_tile(p.g, p.eg, blockIdx.x);
    const bool two = et.two != 0;                      // the tile holds rows of two images (flat tiling only)
    const uint32_t smem_base = smem_u32(smem);
    const bool dec_mode = p.dec_mode != 0;

    const unsigned long long cta_t0 = (p.dbg_acc && tid == 0) ? global_ns() : 0ull;
    const uint32_t tmem = cta_setup(bars, WARP_PRODUCER);
    const uint32_t T = tmem, XA = tmem + 256;

    // source image of the tile's first image (the second one's is src_img + 1) and its length
    int src_img = et.set * p.g.B + et.b0, src_len = et.L;
    if (p.cross) { src_img = et.set == 0 ? p.g.B + et.b0 : et.b0; src_len = et.set == 0 ? p.g.L2 : p.g.L1; }

    if (warp == WARP_PRODUCER) {
        // ------------------------------------------------------------------ residual stream + weight stream
        if (lane == 0) {
            if (!p.load_feat) {   // the tile's residual stream [64 quads][128 rows][4] fp32 = 128 KB, into the (still free) image area
                mbar_arrive_expect_tx(&bars->x_full, 2 * IMG_BYTES);
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.xt + xt_off(blockIdx.x, 0, 0));
#pragma unroll 1
                for (int i = 0; i < 4; ++i)
                    bulk_g2s(smem + SM_AHI + i * (IMG_BYTES / 2), src + (size_t)i * (IMG_BYTES / 2), IMG_BYTES / 2, &bars->x_full);
            }
#pragma unroll 1
            for (int k = 0; k < 3; ++k)
                for (uint32_t off = blockIdx.x * STAGE_BYTES; off < p.pf_bytes[k]; off += gridDim.x * STAGE_BYTES)
                    bulk_prefetch_l2(static_cast<const uint8_t*>(p.pf_ptr[k]) + off, min(STAGE_BYTES, p.pf_bytes[k] - off));
            uint32_t g = 0;
            auto stream = [&](const __half* src, int nunits, int stride) { ring_stream(smem, bars, p.flag, g, src, nunits, stride); };
            if (p.do_q) {
                stream(p.w_q, 4, 4);                                                  // hi units only
                stream(p.mimg + (size_t)src_img * GEMM_HALFS, 8, 2);
                if (two) stream(p.mimg + (size_t)(src_img + 1) * GEMM_HALFS, 8, 2);
                stream(p.w_mlp, 32, 2);
            }
            if (p.do_kv) {
                stream(p.w_kv, 8, 2);                                                 // Wv: hi and lo
                stream(p.w_kv + GEMM_HALFS, 4, 4);                                    // Wk: hi only
            }
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        // ------------------------------------------------------------------ MMA issue
        // TMEM: T = columns [0,256): working accumulator (q, h_a, h_b; v in the source phase);
        //       XA = columns [256,512): THE RESIDUAL STREAM of the tile.  The merge and W2 GEMMs accumulate onto it in
        //       place (x += msg, x += y are done by the tensor core); in the source phase x is dead and XA takes k.
        if (lane == 0) {
            MmaState ms;
            const long long t_begin = clock64();
            const unsigned long long ns_begin = p.dbg_acc ? global_ns() : 0ull;
            auto wait_a = [&](int pass) { mma_wait_a(bars, p.flag, ms, pass); };
            auto gemm = [&](uint32_t d, bool accumulate, bool wait, int terms) { gemm_issue(smem_base, bars, p.flag, ms, d, accumulate, wait, false, terms); };
            if (p.do_q) {
                gemm(T, false, true, T_HH);    umma_commit(&bars->s_full[0]);     // q   = (LNq(x)+pos) Wq^T
                gemm(XA, true, true, T_ALL);   umma_commit(&bars->s_full[1]);     // x  += (phi(q)/Z) M_img^T
                if (two) { gemm(XA, true, true, T_ALL); umma_commit(&bars->s_full[1]); }   // rows of the second image, its M_img
                gemm(T, false, true, T_ALL);   umma_commit(&bars->s_full[0]);     // h_a = LN2(x) W1a^T
                {   // h_b overwrites T: every row thread must have read h_a
                    const long long t0 = clock64();
                    mbar_wait(&bars->s_free, 0, p.flag);
                    ms.t_a += clock64() - t0;
                    tc_fence_after();
                }
                gemm(T, false, false, T_ALL);  umma_commit(&bars->s_full[0]);     // h_b = LN2(x) W1b^T
                gemm(XA, true, true, T_ALL);   umma_commit(&bars->s_full[1]);     // x  += gelu(h_a) W2a^T
                gemm(XA, true, true, T_ALL);   umma_commit(&bars->s_full[1]);     // x  += gelu(h_b) W2b^T
            }
            if (p.do_kv) {
                if (!dec_mode) {
                    gemm(T, false, true, T_ALL);            umma_commit(&bars->s_full[0]);   // v
                    gemm(XA, false, false, T_HH | T_LH);    umma_commit(&bars->s_full[1]);   // k (same image)
                } else {
                    gemm(T, false, true, T_HH | T_HL);      umma_commit(&bars->s_full[0]);   // v = x Wv^T
                    gemm(XA, false, true, T_HH);            umma_commit(&bars->s_full[1]);   // k = (x+pos) Wk^T
                }
                // per 128-channel half: KV = Kf^T V (diagonal 32x32 blocks are the heads); Ksum is reduced by the row warps
                // the token rows are the K dimension, 16 per MMA: a two-image tile splits the k-steps at the image
                // boundary (a multiple of 16 rows) and accumulates the second image's product in XA's columns
                const int ksplit = two ? et.split / 16 : TILE / 16;
                const int nterms = dec_mode ? 1 : 3;
                for (int half = 0; half < 2; ++half) {
                    wait_a(half);
                    const uint32_t kf_hi = smem_base + SM_AHI + KF_OFF, kf_lo = smem_base + SM_ALO + KF_OFF;
                    const uint32_t v_hi = smem_base + SM_AHI + V_OFF, v_lo = smem_base + SM_ALO + V_OFF;
                    for (int term = 0; term < nterms; ++term) {
                        const uint32_t a = term == 1 ? kf_lo : kf_hi, bb = term == 2 ? v_lo : v_hi;
#pragma unroll
                        for (int k = 0; k < TILE / 16; ++k) {
                            const bool second = k >= ksplit;
                            const uint32_t dkv = (second ? XA : T) + half * 128;
                            const uint32_t first_of_group = (term == 0 && (k == 0 || k == ksplit)) ? 0u : 1u;
                            umma_f16(dkv, umma_desc(a + k * 2048, SLAB_BYTES, ATOM_BYTES),
                                     umma_desc(bb + k * 2048, SLAB_BYTES, ATOM_BYTES), IDESC_KV, first_of_group);
                        }
                    }
                    umma_commit(&bars->s_full[half]);
                }
            }
            if (p.dbg_acc && p.do_q && p.do_kv && !p.dec_mode) {
                atomicAdd(p.dbg_acc + DBG_MMA_TOTAL, (unsigned long long)(clock64() - t_begin));
                atomicAdd(p.dbg_acc + DBG_MMA_WAIT_A, (unsigned long long)ms.t_a);
                atomicAdd(p.dbg_acc + DBG_MMA_WAIT_W, (unsigned long long)ms.t_ring);
                atomicAdd(p.dbg_acc + DBG_TILES, 1ull);
                atomicAdd(p.dbg_acc + DBG_TILE_NS, global_ns() - ns_begin);
            }
        }
        __syncwarp();
    } else if (warp < WARP_PRODUCER) {
        // ------------------------------------------------------------------ row warps
        // thread <-> (token row r, column chunks [32*cq, +32) and [128 + 32*cq, +32)).  At most 64 tile values live in a
        // thread's registers at any time (the residual stream lives in TMEM), so the 96-register budget holds.
        const int q = warp & 3, cq = warp >> 2;            // TMEM lane quarter, column quarter
        const int r = q * 32 + lane;                       // token row of the tile
        // which image / token this row is (enc_tile): rows >= split belong to the tile's second image
        const int rel = r >= et.split ? 1 : 0;
        const int rb = et.b0 + rel;                        // image inside the set
        const int rl = rel ? r - et.split : et.l0 + r;     // token inside the image
        const bool valid = rb < et.B && rl < et.L;
        const int pl = valid ? rl : 0;                     // row of the position table
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float* X = reinterpret_cast<float*>(smem + SM_X);       // 512-float scratch: Ksum of the source image(s)
        uint8_t* img_hi = smem + SM_AHI;
        uint8_t* img_lo = smem + SM_ALO;
        const float* post = (et.set == 0 ? p.post1 : p.post2);
        const float* maskp = (et.set == 0 ? p.mask1 : p.mask2);
        const bool has_mask = maskp != nullptr;
        const float mrow = (has_mask && valid) ? __ldg(maskp + (size_t)rb * et.L + rl) : 1.f;
        uint32_t ns0 = 0, ns1 = 0;
        long long t_prev = clock64();
        auto stamp = [&](int i) {                         // OETR_TIMING=1: stage durations of row-warp thread 0
            if (p.dbg_acc && p.do_q && p.do_kv && !p.dec_mode && tid == 0) {
                const long long t = clock64();
                atomicAdd(p.dbg_acc + DBG_STAGE0 + i, (unsigned long long)(t - t_prev));
                t_prev = t;
            }
        };
        auto wait_s = [&](int b) {
            mbar_wait(&bars->s_full[b], (b ? ns1++ : ns0++) & 1, p.flag);
            tc_fence_after();
        };
        auto publish = [&](int pass) {                     // operand image pass written; accumulator reads done
            tc_fence_before();
            fence_async_smem();
            mbar_arrive(&bars->a_full[pass]);
        };
        // the thread's 64 columns of an accumulator (both column passes), one wait
        auto load_acc = [&](uint32_t acc, float (&v)[2][32]) {
            tmem_ld32x2(acc + lane_addr + cq * 32, v[0], v[1]);
        };
        // Positional rows of this token (tile-blocked table: coalesced per warp).  PositionEncodingSine's frequencies are
        // exp(-2k), k = channel / 4 (models/utils.py:188-190 with its precedence quirk): from channel 64 on the angles
        // are below max_w * e^-32 < 4e-11, i.e. the rows are (sin, cos, sin, cos) = (0, 1, 0, 1) to far below fp32
        // resolution of the sum LN(x) + pos.  Only the two chunks with channels < 64 read the table (pass 0 of the
        // column quarters 0 and 1); everywhere else the constant pattern is added: 1/8 of the position-row traffic.
        const bool pos_chunk = cq < 2;                     // warp-uniform
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
        // LayerNorm statistics of the row.  Each of the row's four threads (one per column quarter, in four different
        // warps of the same TMEM lane quarter) reduces its 64 values exactly (local mean, local M2); the (mean, M2) pairs
        // are exchanged through two columns per thread of accumulator T -- columns of the thread's own chunk, which it
        // has consumed -- and merged with Chan's formula.  One 128-thread barrier per LayerNorm.
        auto row_stats = [&](const float (&x)[2][32], float& mean, float& rstd) {
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
        // operand image <- LN(x) [+ pos], both column passes, computed in place in x (x is dead afterwards);
        // split: (hi, lo) image, else one fp16 value per element
        auto ln_image = [&](float (&x)[2][32], const float* __restrict__ gamma, const float* __restrict__ beta, bool with_pos, bool split) {
            float4 ps[8];
            if (with_pos) load_pos(ps);                    // pass 0's rows, in flight during the statistics
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
                    // (x - mean) * rstd * g + b + pos  ==  fma(fma(x, rstd, shift), g, b + pos)
                    x[pass][jq * 4 + 0] = fmaf(fmaf(x[pass][jq * 4 + 0], rstd, shift), gamma[c + 0], beta[c + 0] + pz.x);
                    x[pass][jq * 4 + 1] = fmaf(fmaf(x[pass][jq * 4 + 1], rstd, shift), gamma[c + 1], beta[c + 1] + pz.y);
                    x[pass][jq * 4 + 2] = fmaf(fmaf(x[pass][jq * 4 + 2], rstd, shift), gamma[c + 2], beta[c + 2] + pz.z);
                    x[pass][jq * 4 + 3] = fmaf(fmaf(x[pass][jq * 4 + 3], rstd, shift), gamma[c + 3], beta[c + 3] + pz.w);
                }
                if (with_pos && pass == 0) {               // pass 1: channels >= 128, the constant pattern
#pragma unroll
                    for (int jq = 0; jq < 8; ++jq) ps[jq] = make_float4(0.f, 1.f, 0.f, 1.f);
                }
                if (split) store_row32_split(img_hi, img_lo, r, c0, x[pass]);
                else store_row32_hi(img_hi, r, c0, x[pass]);
                publish(pass);
            }
        };
        // operand image <- x [+ pos] (decoder K/V projections: no LayerNorm), one fp16 value per element; x is kept
        auto raw_image = [&](const float (&x)[2][32], bool with_pos) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;
                float4 ps[8];
                if (with_pos) {
                    if (pass == 0) load_pos(ps);
                    else {
#pragma unroll
                        for (int jq = 0; jq < 8; ++jq) ps[jq] = make_float4(0.f, 1.f, 0.f, 1.f);
                    }
                }
                float v[32];
#pragma unroll
                for (int jq = 0; jq < 8; ++jq) {
                    const float4 pz = with_pos ? ps[jq] : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[jq * 4 + 0] = x[pass][jq * 4 + 0] + pz.x; v[jq * 4 + 1] = x[pass][jq * 4 + 1] + pz.y;
                    v[jq * 4 + 2] = x[pass][jq * 4 + 2] + pz.z; v[jq * 4 + 3] = x[pass][jq * 4 + 3] + pz.w;
                }
                store_row32_hi(img_hi, r, c0, v);
                publish(pass);
            }
        };
        auto store_x = [&](const float (&x)[2][32]) {      // tile-blocked residual stream for the next launch
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;
#pragma unroll
                for (int jq = 0; jq < 8; ++jq)
                    *reinterpret_cast<float4*>(p.xt + xt_off(blockIdx.x, (c0 >> 2) + jq, r)) =
                        make_float4(x[pass][jq * 4], x[pass][jq * 4 + 1], x[pass][jq * 4 + 2], x[pass][jq * 4 + 3]);
            }
        };

        float x[2][32];                                    // the one 64-value register tile of this thread
        // ---- the residual stream of this thread's row
        if (p.load_feat) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;
                const float* feat = et.set == 0 ? p.feat1 : p.feat2;
                const float* f = feat + ((size_t)(valid ? rb : 0) * C + c0) * et.L + pl;
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] = valid ? f[(size_t)e * et.L] : 0.f;
            }
        } else {
            mbar_wait(&bars->x_full, 0, p.flag);           // the bulk copy has landed (complete_tx makes it visible)
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int quad0 = pass * 32 + cq * 8;
#pragma unroll
                for (int jq = 0; jq < 8; ++jq) {
                    const float4 v = *reinterpret_cast<const float4*>(smem + SM_AHI + ((size_t)(quad0 + jq) * TILE + r) * 16);
                    x[pass][jq * 4 + 0] = v.x; x[pass][jq * 4 + 1] = v.y; x[pass][jq * 4 + 2] = v.z; x[pass][jq * 4 + 3] = v.w;
                }
            }
            named_bar_sync(1, N_ROW_THREADS);              // every thread has its rows: the image area may be overwritten
        }
        stamp(0);
        if (p.do_q) {
            // the residual stream moves into TMEM (XA) for the whole query phase
            tmem_st32(XA + lane_addr + cq * 32, x[0]);
            tmem_st32(XA + lane_addr + 128 + cq * 32, x[1]);
            tmem_st_wait();
            // Ksum of the source image(s) -> X (nobody else uses X during the query phase)
            if (tid < 256 || two) X[tid] = __ldg(p.ksum + (size_t)(src_img + (tid >> 8)) * C + (tid & 255));   // [256, 512): second image
            // (E0) A = LNq(x) + pos   (one fp16 value per element: the q GEMM is a 1-term product)
            ln_image(x, p.lnq_g, p.lnq_b, true, false);
            stamp(1);
            named_bar_sync(1, N_ROW_THREADS);              // Ksum visible to every row thread
            const float* Xk = X + rel * 256;
            // (E1) A = phi(q) / Z   (linear_attention.py:33,46; the KV product is folded into M_img)
            wait_s(0);
            stamp(2);
            load_acc(T, x);
            const float eps_s = ATTN_EPS / (float)src_len;    // summaries arrive scaled by 1/S (k_fold)
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;      // one head per 32-column chunk
                float den = 0.f;
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(Xk + c0 + e4 * 4);
                    x[pass][e4 * 4 + 0] = elu1(x[pass][e4 * 4 + 0]); den = fmaf(x[pass][e4 * 4 + 0], k4.x, den);
                    x[pass][e4 * 4 + 1] = elu1(x[pass][e4 * 4 + 1]); den = fmaf(x[pass][e4 * 4 + 1], k4.y, den);
                    x[pass][e4 * 4 + 2] = elu1(x[pass][e4 * 4 + 2]); den = fmaf(x[pass][e4 * 4 + 2], k4.z, den);
                    x[pass][e4 * 4 + 3] = elu1(x[pass][e4 * 4 + 3]); den = fmaf(x[pass][e4 * 4 + 3], k4.w, den);
                }
                // Q = phi(q) * q_mask before Z (linear_attention.py:37,46): phi(q) m / (m den + eps)
                const float inv = mrow / (den * mrow + eps_s);
                // two-image tile: the GEMM with the first image's M_img must see zero rows for the second image's tokens
                const float sc = (two && rel) ? 0.f : inv;
                float v[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) { v[e] = x[pass][e] * sc; x[pass][e] *= inv; }
                store_row32_split(img_hi, img_lo, r, c0, v);
                publish(pass);
            }
            if (two) {                                    // ... and the GEMM with the second image's M_img zero rows for the first's
                wait_s(1);                                // the first GEMM has consumed the image
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    float v[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = rel ? x[pass][e] : 0.f;
                    store_row32_split(img_hi, img_lo, r, pass * 128 + cq * 32, v);
                    publish(pass);
                }
            }
            stamp(3);
            // (E2) A = LN2(x), x = x + msg as accumulated by the tensor core
            wait_s(1);
            stamp(4);
            load_acc(XA, x);
            ln_image(x, p.ln2_g, p.ln2_b, false, true);
            stamp(5);
            // (E3) gelu(h_a): T is read at once (h_b may then overwrite it), the values wait in registers until the GEMMs
            // reading the LN2 image have completed (h_b commit); then h_b is read BEFORE the image is published, because
            // the GEMM it feeds is followed by nothing that protects T ... and x += gelu(h_a) W2a^T runs under (E4)
            wait_s(0);
            stamp(6);
            load_acc(T, x);
            tc_fence_before();
            mbar_arrive(&bars->s_free);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass)
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] = gelu_erf(x[pass][e]);
            wait_s(0);                                     // h_b complete: the LN2 image is free, T holds h_b
            store_row32_split(img_hi, img_lo, r, cq * 32, x[0]);
            store_row32_split(img_hi, img_lo, r, 128 + cq * 32, x[1]);
            load_acc(T, x);
            publish(0);
            publish(1);
            stamp(7);
            // (E4) gelu(h_b) under the W2a GEMM; its image is written once that GEMM has completed
#pragma unroll
            for (int pass = 0; pass < 2; ++pass)
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] = gelu_erf(x[pass][e]);
            wait_s(1);
            store_row32_split(img_hi, img_lo, r, cq * 32, x[0]);
            publish(0);
            store_row32_split(img_hi, img_lo, r, 128 + cq * 32, x[1]);
            publish(1);
            stamp(8);
            // (E5) x = x + y as accumulated by the tensor core
            wait_s(1);
            stamp(9);
            load_acc(XA, x);
            stamp(10);
        }
        if (p.store_x) store_x(x);
        stamp(11);
        if (p.do_kv) {
            if (!dec_mode) {
                ln_image(x, p.lnkv_g, p.lnkv_b, true, true);   // k and v share LN_kv(x)+pos (transformer.py:119-126)
                stamp(12);
                wait_s(0);
                wait_s(1);
            } else {
                raw_image(x, false);                           // v = x Wv^T + bv      (transformer.py:243-249)
                wait_s(0);
                raw_image(x, true);                            // k = (x+pos) Wk^T + bk
                wait_s(1);
            }
            stamp(13);
            // half images (tokens = K dimension): V and Kf = elu(k)+1; padded rows are zero
            // partial summaries of this tile: one slot per image of the tile when the tiling is flat
            float* part = p.kv_part + (size_t)blockIdx.x * (p.eg.flat ? 2 : 1) * PART_FLOATS;
            const float* bvp = p.lnkv_g;                       // dec_mode: bv | bk ride in the lnkv slots
            const float* bkp = p.lnkv_b;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32, ch = cq * 32;      // ch: column inside the 128-channel half
                float v[32];
                tmem_ld32(T + lane_addr + c0, v);
                if (dec_mode) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] += bvp[c0 + e];
                }
                if (!valid) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = 0.f;
                } else if (has_mask) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] *= mrow;
                }
                if (pass == 1) wait_s(0);                               // KV of half 0 has consumed the images
                if (dec_mode) store_row32_hi(img_hi + V_OFF, r, ch, v);
                else store_row32_split(img_hi + V_OFF, img_lo + V_OFF, r, ch, v);
                tmem_ld32(XA + lane_addr + c0, v);
                if (dec_mode) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] += bkp[c0 + e];
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = valid ? elu1(v[e]) * mrow : 0.f;
                if (dec_mode) store_row32_hi(img_hi + KF_OFF, r, ch, v);
                else store_row32_split(img_hi + KF_OFF, img_lo + KF_OFF, r, ch, v);
                publish(pass);
                // Ksum[c0 + j] = sum over the tile's rows of Kf[:, c0 + j] (fp32, exact operands): butterfly
                // transpose-reduce inside the warp (lane j ends with column j summed over the warp's 32 rows); the four
                // row quarters write their partial sums to four slots that k_fold / k_sum_partials add in fixed order
                // (no barrier, deterministic); per image of the tile (rows of the other image masked)
#pragma unroll 1
                for (int im = 0; im < (two ? 2 : 1); ++im) {
                    float w[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) w[e] = (!two || rel == im) ? v[e] : 0.f;
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool up = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < off; ++i) {
                            const float send = up ? w[i] : w[i + off];
                            const float keep = up ? w[i + off] : w[i];
                            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    part[im * PART_FLOATS + NH * HD * HD + q * C + c0 + lane] = w[0];
                }
            }
            stamp(14);
            // results: KV diagonal blocks (this warp's TMEM lanes are the d-channels of head 4*half + q)
            wait_s(1);
            stamp(15);
            if (cq < 2) {
                const int half = cq, h = half * 4 + q;
                for (int im = 0; im < (two ? 2 : 1); ++im) {
                    float v[32];
                    tmem_ld32((im ? XA : T) + lane_addr + half * 128 + q * 32, v);
                    float* o = part + im * PART_FLOATS + h * HD * HD + lane * HD;
#pragma unroll
                    for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(o + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                }
            }
            tc_fence_before();
            stamp(16);
        }
    }
    // teardown
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_PRODUCER) tmem_dealloc(tmem, 512);
    if (tid == 0) dbg_log_cta(p.dbg_acc, p.do_q ? 2 : (p.dec_mode ? 3 : 1), cta_t0);
}

}  // namespace oetr
