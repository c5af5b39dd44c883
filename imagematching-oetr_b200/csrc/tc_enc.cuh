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
    const float *lnq_g, *lnq_b, *ln2_g, *ln2_b;
    const __half* w_q;          // Wq                         (16 stages)
    const __half* w_mlp;        // W1a | W1b | W2a | W2b      (64 stages)
    const __half* mimg;         // [2B images][GEMM_HALFS] folded merge weights of the source image
    const float* ksum;          // [2B images][256]
    int cross;                  // 1: the source is the partner image (transformer.py:354-358)
    // kv phase (encoder layer i+1, or a decoder layer's cross-attention when lnkv_g == nullptr)
    const float *lnkv_g, *lnkv_b;   // nullptr: decoder mode: k = (x+pos) Wk^T + bk, v = x Wv^T + bv
    const float *bk, *bv;
    const __half* w_kv;         // Wv | Wk                    (32 stages)
    float* kv_part;             // [tiles][KVS] per-tile partial summaries
    int* flag;
    unsigned long long* dbg_acc;   // nullable: global cycle accumulators (OETR_TIMING=1), see DBG_* in tc_tiles.cuh
    // L2 prefetch: every layer's weights are read once per forward, so without it each stage is a DRAM-latency
    // miss for the whole first wave.  The grid spreads these ranges (the NEXT launch's weights) in 16 KB pieces.
    const void* pf_ptr[3];
    uint32_t pf_bytes[3];
};

__global__ void __launch_bounds__(N_THREADS, 1) k_enc(const EncParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const EncTile et = enc_tile(p.g, p.eg, blockIdx.x);
    const bool two = et.two != 0;                      // the tile holds rows of two images (flat tiling only)
    const uint32_t smem_base = smem_u32(smem);
    const bool dec_mode = p.lnkv_g == nullptr;

    const uint32_t tmem = cta_setup(bars, WARP_PRODUCER);
    const uint32_t S0 = tmem, S1 = tmem + 256;

    // source image of the tile's first image (the second one's is src_img + 1) and its length
    int src_img = et.set * p.g.B + et.b0, src_len = et.L;
    if (p.cross) { src_img = et.set == 0 ? p.g.B + et.b0 : et.b0; src_len = et.set == 0 ? p.g.L2 : p.g.L1; }

    if (warp == WARP_PRODUCER) {
        // ------------------------------------------------------------------ weight stream
        if (lane == 0) {
#pragma unroll 1
            for (int k = 0; k < 3; ++k)
                for (uint32_t off = blockIdx.x * STAGE_BYTES; off < p.pf_bytes[k]; off += gridDim.x * STAGE_BYTES)
                    bulk_prefetch_l2(static_cast<const uint8_t*>(p.pf_ptr[k]) + off, min(STAGE_BYTES, p.pf_bytes[k] - off));
            uint32_t g = 0;
            auto stream = [&](const __half* src, int nstages) { ring_stream(smem, bars, p.flag, g, src, nstages); };
            if (p.do_q) {
                stream(p.w_q, GEMM_STAGES);
                stream(p.mimg + (size_t)src_img * GEMM_HALFS, GEMM_STAGES);
                if (two) stream(p.mimg + (size_t)(src_img + 1) * GEMM_HALFS, GEMM_STAGES);
                stream(p.w_mlp, 4 * GEMM_STAGES);
            }
            if (p.do_kv) stream(p.w_kv, 2 * GEMM_STAGES);
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        // ------------------------------------------------------------------ MMA issue
        if (lane == 0) {
            MmaState ms;
            const long long t_begin = clock64();
            auto wait_a = [&](int pass) { mma_wait_a(bars, p.flag, ms, pass); };
            auto gemm = [&](uint32_t d, bool accumulate, bool wait) { gemm_issue(smem_base, bars, p.flag, ms, d, accumulate, wait, false); };
            if (p.do_q) {
                gemm(S0, false, true);  umma_commit(&bars->s_full[0]);     // q   = (LNq(x)+pos) Wq^T
                gemm(S1, false, true);  umma_commit(&bars->s_full[1]);     // msg = (phi(q)/Z) M_img^T
                if (two) { gemm(S0, false, false); umma_commit(&bars->s_full[0]); }   // ... with the second image's M_img
                gemm(S0, false, true);  umma_commit(&bars->s_full[0]);     // h_a = LN2(x) W1a^T
                gemm(S1, false, false); umma_commit(&bars->s_full[1]);     // h_b = LN2(x) W1b^T
                gemm(S0, false, true);  umma_commit(&bars->s_full[0]);     // y   = gelu(h_a) W2a^T
                gemm(S0, true, true);   umma_commit(&bars->s_full[0]);     // y  += gelu(h_b) W2b^T
            }
            if (p.do_kv) {
                gemm(S0, false, true);      umma_commit(&bars->s_full[0]); // v
                gemm(S1, false, dec_mode);  umma_commit(&bars->s_full[1]); // k (decoder: from a second image)
                // per 128-channel half: KV = Kf^T V (diagonal 32x32 blocks are the heads); Ksum is reduced by the row warps
                // the token rows are the K dimension, 16 per MMA: a two-image tile splits the k-steps at the image
                // boundary (a multiple of 16 rows) and accumulates the second image's product in S1's columns
                const int ksplit = two ? et.split / 16 : TILE / 16;
                for (int half = 0; half < 2; ++half) {
                    wait_a(half);
                    const uint32_t kf_hi = smem_base + SM_AHI + KF_OFF, kf_lo = smem_base + SM_ALO + KF_OFF;
                    const uint32_t v_hi = smem_base + SM_AHI + V_OFF, v_lo = smem_base + SM_ALO + V_OFF;
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t a = term == 1 ? kf_lo : kf_hi, bb = term == 2 ? v_lo : v_hi;
#pragma unroll
                        for (int k = 0; k < TILE / 16; ++k) {
                            const bool second = k >= ksplit;
                            const uint32_t dkv = (second ? S1 : S0) + half * 128;
                            const uint32_t first_of_group = (term == 0 && (k == 0 || k == ksplit)) ? 0u : 1u;
                            umma_f16(dkv, umma_desc(a + k * 2048, SLAB_BYTES, ATOM_BYTES),
                                     umma_desc(bb + k * 2048, SLAB_BYTES, ATOM_BYTES), IDESC_KV, first_of_group);
                        }
                    }
                    umma_commit(&bars->s_full[half]);
                }
            }
            if (p.dbg_acc) {
                atomicAdd(p.dbg_acc + DBG_MMA_TOTAL, (unsigned long long)(clock64() - t_begin));
                atomicAdd(p.dbg_acc + DBG_MMA_WAIT_A, (unsigned long long)ms.t_a);
                atomicAdd(p.dbg_acc + DBG_MMA_WAIT_W, (unsigned long long)ms.t_ring);
                atomicAdd(p.dbg_acc + DBG_TILES, 1ull);
            }
        }
        __syncwarp();
    } else if (warp < WARP_PRODUCER) {
        // ------------------------------------------------------------------ row warps
        const int q = warp & 3, cq = warp >> 2;            // TMEM lane quarter, column quarter
        const int r = q * 32 + lane;                       // token row of the tile
        // which image / token this row is (enc_tile): rows >= split belong to the tile's second image
        const int rel = r >= et.split ? 1 : 0;
        const int rb = et.b0 + rel;                        // image inside the set
        const int rl = rel ? r - et.split : et.l0 + r;     // token inside the image
        const bool valid = rb < et.B && rl < et.L;
        const int pl = valid ? rl : 0;                     // row of the position table
        const bool warp_has_rel1 = two && et.split < q * 32 + 32;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float* X = reinterpret_cast<float*>(smem + SM_X);       // 512-float scratch, see the shared-memory map
        uint8_t* img_hi = smem + SM_AHI;
        uint8_t* img_lo = smem + SM_ALO;
        const float* post = (et.set == 0 ? p.post1 : p.post2);
        const float* maskp = (et.set == 0 ? p.mask1 : p.mask2);
        const bool has_mask = maskp != nullptr;
        const float mrow = (has_mask && valid) ? __ldg(maskp + (size_t)rb * et.L + rl) : 1.f;
        uint32_t ns0 = 0, ns1 = 0;
        long long t_prev = clock64();
        auto stamp = [&](int i) {                         // OETR_TIMING=1: stage durations of row-warp thread 0
            if (p.dbg_acc && tid == 0) {
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
        // ---- the residual stream of this thread: columns [32*cq, +32) and [128 + 32*cq, +32) of row r
        float x[2][32];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int c0 = pass * 128 + cq * 32;
            if (p.load_feat) {
                const float* feat = et.set == 0 ? p.feat1 : p.feat2;
                const float* f = feat + ((size_t)(valid ? rb : 0) * C + c0) * et.L + pl;
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] = valid ? f[(size_t)e * et.L] : 0.f;
            } else {
#pragma unroll
                for (int jq = 0; jq < 8; ++jq) {
                    const float4 v = *reinterpret_cast<const float4*>(p.xt + xt_off(blockIdx.x, (c0 >> 2) + jq, r));
                    x[pass][jq * 4 + 0] = v.x; x[pass][jq * 4 + 1] = v.y; x[pass][jq * 4 + 2] = v.z; x[pass][jq * 4 + 3] = v.w;
                }
            }
        }
        // two-pass LayerNorm statistics of the row (4 threads per row, combined through X), then (gamma | beta) are
        // staged into X for the normalisation pass (their global loads are issued before the statistics)
        auto ln_stats = [&](const float* __restrict__ gamma, const float* __restrict__ beta, float& mean, float& rstd) {
            const float gb = tid < 256 ? __ldg(gamma + tid) : __ldg(beta + tid - 256);
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) s += x[0][e] + x[1][e];
            X[cq * TILE + r] = s;
            named_bar_sync(1, N_ROW_THREADS);
            mean = (X[0 * TILE + r] + X[1 * TILE + r] + X[2 * TILE + r] + X[3 * TILE + r]) * (1.f / C);
            float sq = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const float d0 = x[0][e] - mean, d1 = x[1][e] - mean;
                sq = fmaf(d0, d0, sq);
                sq = fmaf(d1, d1, sq);
            }
            named_bar_sync(1, N_ROW_THREADS);              // every thread has read the sums
            X[cq * TILE + r] = sq;
            named_bar_sync(1, N_ROW_THREADS);
            const float var = (X[0 * TILE + r] + X[1 * TILE + r] + X[2 * TILE + r] + X[3 * TILE + r]) * (1.f / C);
            rstd = rsqrtf(var + LN_EPS);
            named_bar_sync(1, N_ROW_THREADS);
            X[tid] = gb;                                   // X[0,256) = gamma, X[256,512) = beta
            named_bar_sync(1, N_ROW_THREADS);
        };
        // operand image <- [LN](x) [+ pos], both column passes (gamma == nullptr: no LayerNorm)
        auto image_from_x = [&](const float* __restrict__ gamma, const float* __restrict__ beta, bool with_pos) {
            float mean = 0.f, rstd = 1.f;
            if (gamma) ln_stats(gamma, beta, mean, rstd);
            const float shift = -mean * rstd;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;
                float v[32];
#pragma unroll
                for (int jq = 0; jq < 8; ++jq) {
                    float4 ps = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (with_pos) ps = *reinterpret_cast<const float4*>(post + xt_off(pl >> 7, (c0 >> 2) + jq, pl & 127));
                    float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gamma) {
                        g4 = *reinterpret_cast<const float4*>(X + c0 + jq * 4);
                        b4 = *reinterpret_cast<const float4*>(X + 256 + c0 + jq * 4);
                    }
                    // (x - mean) * rstd * g + b + pos  ==  fma(fma(x, rstd, shift), g, b + pos)
                    v[jq * 4 + 0] = fmaf(fmaf(x[pass][jq * 4 + 0], rstd, shift), g4.x, b4.x + ps.x);
                    v[jq * 4 + 1] = fmaf(fmaf(x[pass][jq * 4 + 1], rstd, shift), g4.y, b4.y + ps.y);
                    v[jq * 4 + 2] = fmaf(fmaf(x[pass][jq * 4 + 2], rstd, shift), g4.z, b4.z + ps.z);
                    v[jq * 4 + 3] = fmaf(fmaf(x[pass][jq * 4 + 3], rstd, shift), g4.w, b4.w + ps.w);
                }
                store_row32_split(img_hi, img_lo, r, c0, v);
                publish(pass);
            }
        };

        stamp(0);
        if (p.do_q) {
            // (E0) A = LNq(x) + pos
            image_from_x(p.lnq_g, p.lnq_b, true);
            stamp(1);
            // Ksum of the source image -> X (every thread is done with gamma/beta after the barrier)
            named_bar_sync(1, N_ROW_THREADS);
            if (tid < 256 || two) X[tid] = __ldg(p.ksum + (size_t)(src_img + (tid >> 8)) * C + (tid & 255));   // [256, 512): second image
            named_bar_sync(1, N_ROW_THREADS);
            const float* Xk = X + rel * 256;
            // (E1) A = phi(q) / Z   (linear_attention.py:33,46; the KV product is folded into M_img)
            wait_s(0);
            stamp(2);
            const float eps_s = ATTN_EPS / (float)src_len;    // summaries arrive scaled by 1/S (k_fold)
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;      // one head per 32-column chunk
                float v[32];
                tmem_ld32(S0 + lane_addr + c0, v);
                float den = 0.f;
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(Xk + c0 + e4 * 4);
                    v[e4 * 4 + 0] = elu1(v[e4 * 4 + 0]); den = fmaf(v[e4 * 4 + 0], k4.x, den);
                    v[e4 * 4 + 1] = elu1(v[e4 * 4 + 1]); den = fmaf(v[e4 * 4 + 1], k4.y, den);
                    v[e4 * 4 + 2] = elu1(v[e4 * 4 + 2]); den = fmaf(v[e4 * 4 + 2], k4.z, den);
                    v[e4 * 4 + 3] = elu1(v[e4 * 4 + 3]); den = fmaf(v[e4 * 4 + 3], k4.w, den);
                }
                if (has_mask) {                           // Q = phi(q) * q_mask, before Z (linear_attention.py:37,46)
                    den *= mrow;
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] *= mrow;
                }
                const float inv = 1.f / (den + eps_s);
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] *= inv;
                store_row32_split(img_hi, img_lo, r, c0, v);
                publish(pass);
            }
            stamp(3);
            // (E2) x += msg ; A = LN2(x)   (two-image tile: rows of the second image take the product with its M_img, S0)
            wait_s(1);
            if (two) wait_s(0);
            stamp(4);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                float v[32];
                tmem_ld32(S1 + lane_addr + pass * 128 + cq * 32, v);
                if (warp_has_rel1) {
                    float v2[32];
                    tmem_ld32(S0 + lane_addr + pass * 128 + cq * 32, v2);
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = rel ? v2[e] : v[e];
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] += v[e];
            }
            image_from_x(p.ln2_g, p.ln2_b, false);
            stamp(5);
            // (E3) A = gelu(h_a): needs h_a (S0) and, for the image to be free, h_b complete (S1)
            // (E4) A = gelu(h_b): the image is free once y = gelu(h_a) W2a^T has completed (S0 commit)
#pragma unroll 1
            for (int which = 0; which < 2; ++which) {
                if (which == 0) { wait_s(0); stamp(6); }
                const uint32_t S = which ? S1 : S0;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const int c0 = pass * 128 + cq * 32;
                    float v[32];
                    tmem_ld32(S + lane_addr + c0, v);
                    // the GEMM that consumes pass 0 overwrites ALL of S0 (h_a): release pass 0 only once this
                    // thread has also read its pass-1 columns
                    if (pass == 1) publish(0);
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = gelu_erf(v[e]);
                    // the image is free once the GEMM still reading it has completed: h_b (S1 commit) before
                    // gelu(h_a) is stored, y = gelu(h_a) W2a^T (S0 commit) before gelu(h_b) is stored
                    if (pass == 0) wait_s(which == 0 ? 1 : 0);
                    store_row32_split(img_hi, img_lo, r, c0, v);
                }
                publish(1);
                stamp(7 + which);
            }
            // (E5) x += y
            wait_s(0);
            stamp(9);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                float v[32];
                tmem_ld32(S0 + lane_addr + pass * 128 + cq * 32, v);
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] += v[e];
            }
            tc_fence_before();
        }
        if (p.store_x) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;
#pragma unroll
                for (int jq = 0; jq < 8; ++jq)
                    *reinterpret_cast<float4*>(p.xt + xt_off(blockIdx.x, (c0 >> 2) + jq, r)) =
                        make_float4(x[pass][jq * 4], x[pass][jq * 4 + 1], x[pass][jq * 4 + 2], x[pass][jq * 4 + 3]);
            }
        }
        stamp(10);
        if (p.do_kv) {
            if (!dec_mode) {
                image_from_x(p.lnkv_g, p.lnkv_b, true);        // k and v share LN_kv(x)+pos (transformer.py:119-126)
                stamp(11);
                wait_s(0);
                wait_s(1);
            } else {
                image_from_x(nullptr, nullptr, false);         // v = x Wv^T + bv      (transformer.py:243-249)
                wait_s(0);
                image_from_x(nullptr, nullptr, true);          // k = (x+pos) Wk^T + bk
                wait_s(1);
            }
            stamp(12);
            // half images (tokens = K dimension): V and Kf = elu(k)+1; padded rows are zero
            // partial summaries of this tile: one slot per image of the tile when the tiling is flat
            float* part = p.kv_part + (size_t)blockIdx.x * (p.eg.flat ? 2 : 1) * KVS;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32, ch = cq * 32;      // ch: column inside the 128-channel half
                float v[32];
                tmem_ld32(S0 + lane_addr + c0, v);
                if (p.bv) {
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bv + c0) + e4);
                        v[e4 * 4] += b4.x; v[e4 * 4 + 1] += b4.y; v[e4 * 4 + 2] += b4.z; v[e4 * 4 + 3] += b4.w;
                    }
                }
                if (!valid) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = 0.f;
                } else if (has_mask) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] *= mrow;
                }
                if (pass == 1) wait_s(0);                               // KV of half 0 has consumed the images
                store_row32_split(img_hi + V_OFF, img_lo + V_OFF, r, ch, v);
                tmem_ld32(S1 + lane_addr + c0, v);
                if (p.bk) {
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bk + c0) + e4);
                        v[e4 * 4] += b4.x; v[e4 * 4 + 1] += b4.y; v[e4 * 4 + 2] += b4.z; v[e4 * 4 + 3] += b4.w;
                    }
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = valid ? elu1(v[e]) * mrow : 0.f;
                store_row32_split(img_hi + KF_OFF, img_lo + KF_OFF, r, ch, v);
                publish(pass);
                // Ksum[c0 + j] = sum over the tile's rows of Kf[:, c0 + j] (fp32, exact operands): butterfly
                // transpose-reduce inside the warp (lane j ends with column j summed over the warp's 32 rows),
                // then across the 4 row quarters through X; per image of the tile (rows of the other image masked)
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
                    X[(cq * 4 + q) * 32 + lane] = w[0];
                    named_bar_sync(1, N_ROW_THREADS);
                    if (q == 0)
                        part[im * KVS + NH * HD * HD + c0 + lane] = X[(cq * 4 + 0) * 32 + lane] + X[(cq * 4 + 1) * 32 + lane] +
                                                                    X[(cq * 4 + 2) * 32 + lane] + X[(cq * 4 + 3) * 32 + lane];
                    named_bar_sync(1, N_ROW_THREADS);
                }
            }
            stamp(13);
            // results: KV diagonal blocks (this warp's TMEM lanes are the d-channels of head 4*half + q)
            wait_s(1);
            stamp(14);
            if (cq < 2) {
                const int half = cq, h = half * 4 + q;
                for (int im = 0; im < (two ? 2 : 1); ++im) {
                    float v[32];
                    tmem_ld32((im ? S1 : S0) + lane_addr + half * 128 + q * 32, v);
                    float* o = part + im * KVS + h * HD * HD + lane * HD;
#pragma unroll
                    for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(o + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                }
            }
            tc_fence_before();
            stamp(15);
        }
    }
    // teardown
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_PRODUCER) tmem_dealloc(tmem, 512);
}

}  // namespace oetr
