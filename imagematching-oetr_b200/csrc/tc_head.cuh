// Decoder and overlap head: k_att, k_conv (tcgen05 3x3 heat-map convolution), k_logits, k_box, k_decoder.
// Part of the tcgen05 (OETR_PREC_FP16) path; compiled into tc_kernels.cu (one translation unit: kernels are
// launched from the host code there).
#pragma once
#include "tc_tiles.cuh"

namespace oetr {
using namespace tc;

// ---------------------------------------------------------------------------------------------------------
// k_att: att[l] = <memory[l,:], hs[img,:]> per token (src/model.py:147-149), tile-blocked like xt; 0 on padding
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_att(const float* __restrict__ xt, TileGeom g, const float* __restrict__ hs,
                                             float* __restrict__ att) {
    __shared__ float part[TILE];
    __shared__ __align__(16) float hsv[C];
    const TileInfo ti = tile_info(g, blockIdx.x);
    const int r = threadIdx.x & 127, half = threadIdx.x >> 7;
    hsv[threadIdx.x] = hs[(size_t)ti.img * C + threadIdx.x];
    __syncthreads();
    float acc = 0.f;
#pragma unroll 8
    for (int jq = 0; jq < 32; ++jq) {
        const int quad = half * 32 + jq;
        const float4 v = *reinterpret_cast<const float4*>(xt + xt_off(blockIdx.x, quad, r));
        const float4 h = *reinterpret_cast<const float4*>(hsv + quad * 4);
        acc += v.x * h.x + v.y * h.y + v.z * h.z + v.w * h.w;
    }
    if (half == 1) part[r] = acc;
    __syncthreads();
    if (half == 0) att[(size_t)blockIdx.x * TILE + r] = r < ti.valid ? acc + part[r] : 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// k_conv: heatmap_conv.0 (3x3, 256->256, pad 1, src/model.py:65-77,152-161) as an implicit GEMM on the tensor
// cores.  heat[l,:] = memory[l,:] * att[l] is formed on the fly; for each of the 9 taps the row warps gather the
// shifted rows (zero outside the map) into the operand image and the MMA warp accumulates
// Y += G_tap . W_tap^T into S0 (3-term split).  The image of tap t+1 is written per column pass as soon as the
// MMAs of tap t that read that pass have completed (a_free), so the gather overlaps the tensor work.
// ---------------------------------------------------------------------------------------------------------
struct ConvParams {
    TileGeom g;
    int hf1, wf1, hf2, wf2;
    const float* xt;            // tile-blocked encoder output (memory)
    const float* att;           // tile-blocked per-token scale
    const __half* w;            // 9 tap GEMM images
    const float* bias;          // heatmap_conv.0.bias
    float* Y;                   // token-major [B*L1 + B*L2][256]
    float* gstat;               // [tiles][32 groups][2]: per-tile GroupNorm partials (mean, M2) over the valid rows
    int* flag;
    unsigned long long* dbg_acc;   // nullable, like EncParams::dbg_acc (slots DBG_CONV..)
};

__global__ void __launch_bounds__(N_THREADS, 1) k_conv(const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const TileInfo ti = tile_info(p.g, blockIdx.x);
    const uint32_t smem_base = smem_u32(smem);
    const unsigned long long cta_t0 = (p.dbg_acc && tid == 0) ? global_ns() : 0ull;
    const uint32_t tmem = cta_setup(bars, WARP_PRODUCER);
    const uint32_t S0 = tmem;

    if (warp == WARP_PRODUCER) {
        if (lane == 0) {
            uint32_t g = 0;
            ring_stream(smem, bars, p.flag, g, p.w, 9 * GEMM_STAGES / 2);
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        if (lane == 0) {
            MmaState ms;
            const long long t_begin = clock64();
            for (int tap = 0; tap < 9; ++tap) gemm_issue(smem_base, bars, p.flag, ms, S0, tap > 0, true, true);
            umma_commit(&bars->s_full[0]);
            if (p.dbg_acc) {
                atomicAdd(p.dbg_acc + DBG_CONV + 0, (unsigned long long)(clock64() - t_begin));
                atomicAdd(p.dbg_acc + DBG_CONV + 1, (unsigned long long)ms.t_a);
                atomicAdd(p.dbg_acc + DBG_CONV + 2, (unsigned long long)ms.t_ring);
                atomicAdd(p.dbg_acc + DBG_CONV + 3, 1ull);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3, cq = warp >> 2;
        const int r = q * 32 + lane;
        const bool valid = r < ti.valid;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int wf = ti.set == 0 ? p.wf1 : p.wf2, hf = ti.set == 0 ? p.hf1 : p.hf2;
        const int l = ti.ti * TILE + r;
        const int y0 = l / wf, x0 = l - y0 * wf;
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            const bool ok = valid && (unsigned)(y0 + dy) < (unsigned)hf && (unsigned)(x0 + dx) < (unsigned)wf;
            const int n = ok ? l + dy * wf + dx : 0;
            const int tile_n = ti.first_tile_of_img + (n >> 7), rn = n & 127;
            const float a = ok ? p.att[(size_t)tile_n * TILE + rn] : 0.f;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = pass * 128 + cq * 32;
                float v[32];
#pragma unroll
                for (int jq = 0; jq < 8; ++jq) {
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) t = *reinterpret_cast<const float4*>(p.xt + xt_off(tile_n, (c0 >> 2) + jq, rn));
                    v[jq * 4 + 0] = t.x * a; v[jq * 4 + 1] = t.y * a; v[jq * 4 + 2] = t.z * a; v[jq * 4 + 3] = t.w * a;
                }
                if (tap > 0) mbar_wait(&bars->a_free[pass], (tap - 1) & 1, p.flag);
                store_row32_split(smem + SM_AHI, smem + SM_ALO, r, c0, v);
                fence_async_smem();
                mbar_arrive(&bars->a_full[pass]);
            }
        }
        mbar_wait(&bars->s_full[0], 0, p.flag);
        tc_fence_after();
        float* red = reinterpret_cast<float*>(smem + SM_X);           // [4 row quarters][32 groups]
        const size_t row = (ti.set == 0 ? (size_t)ti.b * p.g.L1 : (size_t)p.g.B * p.g.L1 + (size_t)ti.b * p.g.L2) + l;
        float y[2][32];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int c0 = pass * 128 + cq * 32;
            tmem_ld32(S0 + lane_addr + c0, y[pass]);
#pragma unroll
            for (int jq = 0; jq < 8; ++jq) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + jq);
                y[pass][jq * 4 + 0] += b4.x; y[pass][jq * 4 + 1] += b4.y; y[pass][jq * 4 + 2] += b4.z; y[pass][jq * 4 + 3] += b4.w;
                if (valid)
                    *reinterpret_cast<float4*>(p.Y + row * C + c0 + jq * 4) =
                        make_float4(y[pass][jq * 4], y[pass][jq * 4 + 1], y[pass][jq * 4 + 2], y[pass][jq * 4 + 3]);
            }
        }
        tc_fence_before();
        // GroupNorm partials of this tile (32 groups of 8 channels, src/model.py:71): exact two-pass (mean, M2) over the
        // valid rows; k_logits merges the tiles of an image with Chan's update.  Thread: groups pass*16 + cq*4 + {0..3}
        const float inv_n = 1.f / (8.f * (float)ti.valid);
        float mean_t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int pass = i >> 2, g4 = i & 3;
            float sgrp = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) sgrp += y[pass][g4 * 8 + e];
            sgrp = valid ? sgrp : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sgrp += __shfl_xor_sync(0xffffffffu, sgrp, o);
            if (lane == 0) red[q * 32 + pass * 16 + cq * 4 + g4] = sgrp;
        }
        named_bar_sync(1, N_ROW_THREADS);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gi = (i >> 2) * 16 + cq * 4 + (i & 3);
            mean_t[i] = (red[gi] + red[32 + gi] + red[64 + gi] + red[96 + gi]) * inv_n;
        }
        named_bar_sync(1, N_ROW_THREADS);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int pass = i >> 2, g4 = i & 3;
            float m2 = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float d = y[pass][g4 * 8 + e] - mean_t[i]; m2 = fmaf(d, d, m2); }
            m2 = valid ? m2 : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
            if (lane == 0) red[q * 32 + pass * 16 + cq * 4 + g4] = m2;
        }
        named_bar_sync(1, N_ROW_THREADS);
        if (q == 0 && lane == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int gi = (i >> 2) * 16 + cq * 4 + (i & 3);
                float* o = p.gstat + ((size_t)blockIdx.x * 32 + gi) * 2;
                o[0] = mean_t[i];
                o[1] = red[gi] + red[32 + gi] + red[64 + gi] + red[96 + gi];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_PRODUCER) tmem_dealloc(tmem, 512);
    if (tid == 0) dbg_log_cta(p.dbg_acc, 4, cta_t0);
}

// ---------------------------------------------------------------------------------------------------------
// k_logits: GroupNorm (image statistics merged from the tile partials) -> ReLU -> 1x1 conv (src/model.py:71-76);
// one CTA per tile, one warp per token.  z is tile-blocked: token l of an image sits at z[first_tile*128 + l].
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_logits(const float* __restrict__ Y, const float* __restrict__ gstat, TileGeom g,
                                                const float* __restrict__ gn_g, const float* __restrict__ gn_b,
                                                const float* __restrict__ w3, const float* __restrict__ b3,
                                                float* __restrict__ z) {
    __shared__ float gm[32], gr[32];
    __shared__ __align__(16) float sc[C], sh[C], w3s[C];
    const TileInfo ti = tile_info(g, blockIdx.x);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid < 32) {
        float n = 0.f, mean = 0.f, m2 = 0.f;
        for (int t = 0; t < ti.T; ++t) {                                   // Chan et al. pairwise update, fixed order
            const float nb = 8.f * (float)min(TILE, ti.L - t * TILE);
            const float* st = gstat + ((size_t)(ti.first_tile_of_img + t) * 32 + tid) * 2;
            const float delta = st[0] - mean, ntot = n + nb;
            mean += delta * nb / ntot;
            m2 += st[1] + delta * delta * n * nb / ntot;
            n = ntot;
        }
        gm[tid] = mean;
        gr[tid] = rsqrtf(m2 / n + GN_EPS);
    }
    __syncthreads();
    {
        const float rs = gr[tid >> 3] * gn_g[tid];
        sc[tid] = rs; sh[tid] = gn_b[tid] - gm[tid >> 3] * rs; w3s[tid] = w3[tid];
    }
    __syncthreads();
    const float bias = b3[0];
    const size_t row0 = (ti.set == 0 ? (size_t)ti.b * g.L1 : (size_t)g.B * g.L1 + (size_t)ti.b * g.L2) + (size_t)ti.ti * TILE;
    float4 s0 = reinterpret_cast<const float4*>(sc)[lane], s1 = reinterpret_cast<const float4*>(sc)[lane + 32];
    float4 h0 = reinterpret_cast<const float4*>(sh)[lane], h1 = reinterpret_cast<const float4*>(sh)[lane + 32];
    float4 w0 = reinterpret_cast<const float4*>(w3s)[lane], w1 = reinterpret_cast<const float4*>(w3s)[lane + 32];
#pragma unroll 4
    for (int r = w; r < ti.valid; r += 8) {
        const float4 a = reinterpret_cast<const float4*>(Y + (row0 + r) * C)[lane];
        const float4 b = reinterpret_cast<const float4*>(Y + (row0 + r) * C)[lane + 32];
        float acc = w0.x * fmaxf(fmaf(a.x, s0.x, h0.x), 0.f);
        acc = fmaf(w0.y, fmaxf(fmaf(a.y, s0.y, h0.y), 0.f), acc);
        acc = fmaf(w0.z, fmaxf(fmaf(a.z, s0.z, h0.z), 0.f), acc);
        acc = fmaf(w0.w, fmaxf(fmaf(a.w, s0.w, h0.w), 0.f), acc);
        acc = fmaf(w1.x, fmaxf(fmaf(b.x, s1.x, h1.x), 0.f), acc);
        acc = fmaf(w1.y, fmaxf(fmaf(b.y, s1.y, h1.y), 0.f), acc);
        acc = fmaf(w1.z, fmaxf(fmaf(b.z, s1.z, h1.z), 0.f), acc);
        acc = fmaf(w1.w, fmaxf(fmaf(b.w, s1.w, h1.w), 0.f), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) z[(size_t)blockIdx.x * TILE + r] = acc + bias;
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_box: softmax over the tokens + soft-argmax on the (x+0.5, y+0.5)*stride grid, stride = img_h / hf for both
// axes (src/model.py:173-184), box assembly from (cx,cy) and tlbr (models/utils.py:16-28 / model.py:193-211)
// ---------------------------------------------------------------------------------------------------------
struct BoxParams {
    TileGeom g;
    int hf1, wf1, hf2, wf2, img_h1, img_w1, img_h2, img_w2, clamp;
    const float* z;             // tile-blocked logits
    const float *mask1, *mask2; // nullable [B][L]: logits where the mask is 0 count as -1e9 (src/model.py:167-171)
    const float* tlbr;          // [2B][4] sigmoid(top,left,bottom,right) from k_decoder
    float *boxes1, *boxes2, *dbg_cxy, *dbg_tlbr;
};
__global__ void __launch_bounds__(256) k_box(const BoxParams p) {
    __shared__ float red[8];
    const int img = blockIdx.x, set = img / p.g.B, b = img % p.g.B;
    const int L = set == 0 ? p.g.L1 : p.g.L2, wf = set == 0 ? p.wf1 : p.wf2, hf = set == 0 ? p.hf1 : p.hf2;
    const int img_h = set == 0 ? p.img_h1 : p.img_h2, img_w = set == 0 ? p.img_w1 : p.img_w2;
    const int first = set == 0 ? b * p.g.T1 : p.g.B * p.g.T1 + b * p.g.T2;
    const float* z = p.z + (size_t)first * TILE;
    const float* mk = set == 0 ? p.mask1 : p.mask2;
    if (mk) mk += (size_t)b * L;
    auto logit = [&](int l) { return (mk && mk[l] == 0.f) ? -1e9f : z[l]; };
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    auto block_reduce = [&](float v, bool is_max) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const float t = __shfl_xor_sync(0xffffffffu, v, o); v = is_max ? fmaxf(v, t) : v + t; }
        __syncthreads();
        if (lane == 0) red[w] = v;
        __syncthreads();
        float r = red[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
        return r;
    };
    float mx = -INFINITY;
    for (int l = tid; l < L; l += 256) mx = fmaxf(mx, logit(l));
    mx = block_reduce(mx, true);
    const float stride = (float)(img_h / hf);
    float se = 0.f, sx = 0.f, sy = 0.f;
    for (int l = tid; l < L; l += 256) {
        const float e = expf(logit(l) - mx);
        se += e;
        sx = fmaf(e, ((float)(l % wf) + 0.5f) * stride, sx);
        sy = fmaf(e, ((float)(l / wf) + 0.5f) * stride, sy);
    }
    se = block_reduce(se, false); sx = block_reduce(sx, false); sy = block_reduce(sy, false);
    if (tid == 0) {
        const float cx = sx / se, cy = sy / se;
        const float* tl = p.tlbr + (size_t)img * 4;
        const float W_ = (float)img_w, H_ = (float)img_h;
        float x1 = cx - tl[1] * W_, y1 = cy - tl[0] * H_, x2 = cx + tl[3] * W_, y2 = cy + tl[2] * H_;
        if (p.clamp) {
            x1 = fminf(fmaxf(x1, 0.f), W_); x2 = fminf(fmaxf(x2, 0.f), W_);
            y1 = fminf(fmaxf(y1, 0.f), H_); y2 = fminf(fmaxf(y2, 0.f), H_);
        }
        float* o = (set == 0 ? p.boxes1 : p.boxes2) + (size_t)b * 4;
        o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
        if (p.dbg_cxy) { p.dbg_cxy[img * 2] = cx; p.dbg_cxy[img * 2 + 1] = cy; }
        if (p.dbg_tlbr) { for (int i = 0; i < 4; ++i) p.dbg_tlbr[img * 4 + i] = tl[i]; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_decoder: the whole 2-layer query decoder (transformer.py:224-284,361-381) for DEC_R query tokens per CTA.
// fp32 on the CUDA cores: thread n owns output channel n of every projection; the transposed weights
// WT[k][n] make the loads coalesced; each weight is read once per CTA and used for DEC_R rows.
// Rows: [0,B) = image set 1 with query_embed1, [B,2B) = set 2 with query_embed2.
// ---------------------------------------------------------------------------------------------------------
constexpr int DEC_R = 2;
struct DecLayerT {
    const float *sa_bq, *sa_bk, *sa_bv, *ca_bq;
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
};
struct DecParams {
    DecLayerT layer[N_DEC];
    const float* qe;            // query_embed1 | query_embed2 (adjacent, [2][256])
    const float* kvs;           // [N_DEC][2B][KVS] cross-attention summaries of the memory
    float* hs;                  // out [2B][256]
    const float* wt;            // all transposed fp32 weights in consumption order (see k_decoder)
    const float *tl_w2, *tl_b2; // tlbr_reg.2: [4][256], [4]  (src/model.py:59-63)
    float* tlbr;                // out [2B][4] sigmoid(top,left,bottom,right)
    int B;
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Weight streaming: the transposed fp32 weights of both layers and tlbr_reg.0 are ONE contiguous array in the
// order the kernel consumes them (sa.wq | sa.wk | sa.wv | sa.wm | ca.wq | ca.wm | w1 | w2 per layer, then tl_w0),
// so a producer lane streams it linearly in 32 KB bulk copies through a 5-stage ring, running ahead across
// matvec boundaries; the 256 compute threads never wait on global-memory latency (the first version, with
// register loads, was bound by the bytes one SM can keep in flight: 185 us for 5.8 MB).
constexpr int DEC_THREADS = 256 + 32;                 // warps 0-7 compute, warp 8 = producer
constexpr int DEC_STAGES = 5;
constexpr uint32_t DEC_CHUNK_BYTES = 32768;
constexpr int DEC_CHUNK_FLOATS = DEC_CHUNK_BYTES / 4;
constexpr uint32_t DEC_SMEM = DEC_STAGES * DEC_CHUNK_BYTES + 256;
struct DecRing { uint64_t* full; uint64_t* empty; const float* stage0; uint32_t g; };

// acc[r][c] = sum_k WT[k][n + 256*c] * xin[r][k]   (c < N/256), WT consumed from the ring (K*N*4/32 KB chunks)
template <int K, int N>
__device__ __forceinline__ void dec_matvec(DecRing& ring, const float* xin, float (&acc)[DEC_R][N / 256]) {
    constexpr int ROWS = DEC_CHUNK_FLOATS / N;        // k rows per chunk
    const int n = threadIdx.x;
#pragma unroll
    for (int r = 0; r < DEC_R; ++r)
#pragma unroll
        for (int c = 0; c < N / 256; ++c) acc[r][c] = 0.f;
    for (int k0 = 0; k0 < K; k0 += ROWS, ++ring.g) {
        const int st = ring.g % DEC_STAGES;
        mbar_wait(&ring.full[st], (ring.g / DEC_STAGES) & 1, nullptr);
        const float* ws = ring.stage0 + (size_t)st * DEC_CHUNK_FLOATS;
#pragma unroll 4
        for (int kk = 0; kk < ROWS; kk += 4) {
            float4 xv[DEC_R];
#pragma unroll
            for (int r = 0; r < DEC_R; ++r) xv[r] = *reinterpret_cast<const float4*>(xin + r * K + k0 + kk);
#pragma unroll
            for (int c = 0; c < N / 256; ++c) {
                const float w0 = ws[(kk + 0) * N + n + 256 * c], w1 = ws[(kk + 1) * N + n + 256 * c];
                const float w2 = ws[(kk + 2) * N + n + 256 * c], w3 = ws[(kk + 3) * N + n + 256 * c];
#pragma unroll
                for (int r = 0; r < DEC_R; ++r)
                    acc[r][c] = fmaf(w3, xv[r].w, fmaf(w2, xv[r].z, fmaf(w1, xv[r].y, fmaf(w0, xv[r].x, acc[r][c]))));
            }
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&ring.empty[st]);
    }
}
// out[r][:] = LN(in[r][:]) (two-pass variance); warps 0..DEC_R-1, one row each
__device__ __forceinline__ void dec_ln(const float* in, const float* __restrict__ g, const float* __restrict__ b, float* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < DEC_R) {
        float v[8], s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] = in[warp * C + lane + 32 * i]; s += v[i]; }
        const float mu = warp_sum_f(s) * (1.f / C);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - mu; sq = fmaf(d, d, sq); }
        const float rstd = rsqrtf(warp_sum_f(sq) * (1.f / C) + LN_EPS);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[warp * C + lane + 32 * i] = (v[i] - mu) * rstd * g[lane + 32 * i] + b[lane + 32 * i];
    }
}

__global__ void __launch_bounds__(DEC_THREADS) k_decoder(const DecParams p) {
    extern __shared__ __align__(1024) uint8_t dsm[];
    __shared__ __align__(16) float t[DEC_R * C], u[DEC_R * C], a[DEC_R * C], qv[DEC_R * C], hid[DEC_R * FF];
    uint64_t* bars = reinterpret_cast<uint64_t*>(dsm + DEC_STAGES * DEC_CHUNK_BYTES);
    const int n = threadIdx.x, lane = n & 31;
    if (n == 0) {
        for (int i = 0; i < DEC_STAGES; ++i) { mbar_init(&bars[i], 1); mbar_init(&bars[DEC_STAGES + i], 8); }
        fence_mbar_init();
    }
    __syncthreads();
    if (n >= 256) {
        // ---- producer: the whole weight array, linearly
        if (lane == 0) {
            const uint32_t total = (uint32_t)((N_DEC * DEC_T_FLOATS + (size_t)C * C) * sizeof(float) / DEC_CHUNK_BYTES);
            for (uint32_t g = 0; g < total; ++g) {
                const int st = g % DEC_STAGES;
                mbar_wait(&bars[DEC_STAGES + st], ((g / DEC_STAGES) & 1) ^ 1, nullptr);
                mbar_arrive_expect_tx(&bars[st], DEC_CHUNK_BYTES);
                bulk_g2s(dsm + (size_t)st * DEC_CHUNK_BYTES, reinterpret_cast<const uint8_t*>(p.wt) + (size_t)g * DEC_CHUNK_BYTES,
                         DEC_CHUNK_BYTES, &bars[st]);
            }
        }
        return;
    }
    DecRing ring{bars, bars + DEC_STAGES, reinterpret_cast<const float*>(dsm), 0u};
    const int row0 = blockIdx.x * DEC_R, rows = 2 * p.B;
    float qe[DEC_R];
#pragma unroll
    for (int r = 0; r < DEC_R; ++r) {
        const int row = min(row0 + r, rows - 1);
        qe[r] = p.qe[(row >= p.B ? C : 0) + n];
        t[r * C + n] = 0.f;                                                    // tgt = zeros (transformer.py:361)
    }
    named_bar_sync(2, 256);
    for (int j = 0; j < N_DEC; ++j) {
        const DecLayerT& w = p.layer[j];
        float acc[DEC_R][1], kk[DEC_R][1], vv[DEC_R][1], h2[DEC_R][2];
        // ---- self-attention over the single query token (transformer.py:236-241, linear_attention.py:22-50)
        dec_ln(t, w.ln1_g, w.ln1_b, u);
        named_bar_sync(2, 256);
#pragma unroll
        for (int r = 0; r < DEC_R; ++r) a[r * C + n] = u[r * C + n] + qe[r];
        named_bar_sync(2, 256);
        dec_matvec<C, C>(ring, a, acc);
        dec_matvec<C, C>(ring, a, kk);
        dec_matvec<C, C>(ring, u, vv);
        named_bar_sync(2, 256);                                                // all reads of a[] done
#pragma unroll
        for (int r = 0; r < DEC_R; ++r) {
            const float qf = elu1(acc[r][0] + w.sa_bq[n]), kf = elu1(kk[r][0] + w.sa_bk[n]);
            const float sden = warp_sum_f(qf * kf);                           // warp = head
            a[r * C + n] = (vv[r][0] + w.sa_bv[n]) * sden / (sden + ATTN_EPS); // KV = kf v^T, Z = 1/(qf.kf + eps)
        }
        named_bar_sync(2, 256);
        dec_matvec<C, C>(ring, a, acc);
#pragma unroll
        for (int r = 0; r < DEC_R; ++r) t[r * C + n] += acc[r][0];
        named_bar_sync(2, 256);
        // ---- cross-attention into the memory summaries (transformer.py:243-250)
        dec_ln(t, w.ln2_g, w.ln2_b, u);
        named_bar_sync(2, 256);
#pragma unroll
        for (int r = 0; r < DEC_R; ++r) a[r * C + n] = u[r * C + n] + qe[r];
        named_bar_sync(2, 256);
        dec_matvec<C, C>(ring, a, acc);
#pragma unroll
        for (int r = 0; r < DEC_R; ++r) qv[r * C + n] = elu1(acc[r][0] + w.ca_bq[n]);
        named_bar_sync(2, 256);
        {
            const int h = n >> 5;
#pragma unroll
            for (int r = 0; r < DEC_R; ++r) {
                const int row = min(row0 + r, rows - 1);
                const float* kv = p.kvs + ((size_t)j * rows + row) * KVS;
                const float den = warp_sum_f(qv[r * C + n] * kv[NH * HD * HD + n]);
                float o = 0.f;
#pragma unroll 8
                for (int d = 0; d < HD; ++d) o = fmaf(qv[r * C + h * HD + d], kv[(h * HD + d) * HD + lane], o);
                a[r * C + n] = o / (den + ATTN_EPS);
            }
        }
        named_bar_sync(2, 256);
        dec_matvec<C, C>(ring, a, acc);
#pragma unroll
        for (int r = 0; r < DEC_R; ++r) t[r * C + n] += acc[r][0];
        named_bar_sync(2, 256);
        // ---- feed-forward (transformer.py:252-254)
        dec_ln(t, w.ln3_g, w.ln3_b, u);
        named_bar_sync(2, 256);
        dec_matvec<C, FF>(ring, u, h2);
#pragma unroll
        for (int r = 0; r < DEC_R; ++r) { hid[r * FF + n] = fmaxf(h2[r][0], 0.f); hid[r * FF + C + n] = fmaxf(h2[r][1], 0.f); }
        named_bar_sync(2, 256);
        dec_matvec<FF, C>(ring, hid, acc);
#pragma unroll
        for (int r = 0; r < DEC_R; ++r) t[r * C + n] += acc[r][0];
        named_bar_sync(2, 256);
    }
#pragma unroll
    for (int r = 0; r < DEC_R; ++r)
        if (row0 + r < rows) p.hs[(size_t)(row0 + r) * C + n] = t[r * C + n];
    // ---- size regression (src/model.py:188-191): sigmoid(W_b relu(W_a hs) + b)
    {
        float acc[DEC_R][1];
        dec_matvec<C, C>(ring, t, acc);
#pragma unroll
        for (int r = 0; r < DEC_R; ++r) a[r * C + n] = fmaxf(acc[r][0], 0.f);
        named_bar_sync(2, 256);
        const int w = n >> 5;
        if (w < 4) {
#pragma unroll
            for (int r = 0; r < DEC_R; ++r) {
                float o = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) o = fmaf(p.tl_w2[(size_t)w * C + lane + 32 * i], a[r * C + lane + 32 * i], o);
                o = warp_sum_f(o);
                if (lane == 0 && row0 + r < rows) p.tlbr[(size_t)(row0 + r) * 4 + w] = 1.f / (1.f + expf(-(o + p.tl_b2[w])));
            }
        }
    }
}

__global__ void k_transpose(const float* __restrict__ W, int N, int K, float* __restrict__ WT) {   // W[N][K] -> WT[K][N]
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) tile[i][threadIdx.x] = W[(size_t)(n0 + i) * K + k0 + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) WT[(size_t)(k0 + i) * N + n0 + threadIdx.x] = tile[threadIdx.x][i];
}

}  // namespace oetr
