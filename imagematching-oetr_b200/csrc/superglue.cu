// SuperGlue's two hot operators on the device, the heaviest consumer downstream of the overlap
// boxes in the reference's pipeline (evaluation.py:125-224 -> third_party/SuperGluePretrainedNetwork/models/superglue.py):
//   k_sg_attention_tc / k_sg_attention   `attention(query, key, value)` (superglue.py:86-90) for the 4 heads x 64 dims of
//                    MultiHeadedAttention (:93-108): softmax(Q K^T / 8) V with an online softmax over key chunks -- the
//                    [N, M] score matrix never exists in memory (the reference materialises it twice per layer, 18 layers).
//                    _tc (default): QK^T and PV on tcgen05 with 3-term split fp16 operands, S and O in TMEM; the second
//                    kernel is the same operator in fp32 on the CUDA cores (mode OETR_SG_FP32; cross-check and baseline).
//   k_sg_sinkhorn / k_sg_transpose / k_sg_finish   `log_optimal_transport` (superglue.py:150-184): log-space Sinkhorn
//                    iterations on the score matrix augmented by the dustbin row / column.  The augmented matrix is never
//                    built (the dustbin entries are the scalar alpha), u and v live in a small workspace, every
//                    half-iteration is one pass over the L2-resident scores with a chunked log-sum-exp (a warp per row, 8
//                    independent loads and exponentials per lane and step); the column update reads a transposed copy made
//                    once per call, so both passes are coalesced.  ALL iterations run inside ONE persistent cooperative
//                    kernel with a grid barrier between half-iterations (one launch instead of 2 x iters).
// Layout: channel-major like the reference's Conv1d tensors, q [batch][256][N] with channel c = d * 4 + h
// (`.view(batch, dim, heads, -1)`, superglue.py:103-104).
#include "../../include/oetr_b200.h"
#include "tc_common.cuh"

#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace {

constexpr int SG_C = 256, SG_H = 4, SG_D = 64;
constexpr int BQ = 64, BK = 64, VP = 68;
constexpr size_t ATT_SMEM = (size_t)(BQ * SG_D + BK * SG_D + BK * VP + BQ * VP) * sizeof(float);     // 67.6 KB

__global__ void __launch_bounds__(256) k_sg_attention(const float* __restrict__ q, const float* __restrict__ k,
                                                      const float* __restrict__ v, float* __restrict__ out, int N, int M) {
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;                      // [d][r]
    float* Ks = Qs + BQ * SG_D;          // [d][c]
    float* Vs = Ks + BK * SG_D;          // [c][d], pitch 68
    float* Ps = Vs + BK * VP;            // [r][c], pitch 68
    const int h = blockIdx.y, b = blockIdx.z, n0 = blockIdx.x * BQ;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const float* qb = q + (size_t)b * SG_C * N;
    const float* kb = k + (size_t)b * SG_C * M;
    const float* vb = v + (size_t)b * SG_C * M;
    for (int i = tid; i < BQ * SG_D; i += 256) {
        const int d = i >> 6, r = i & 63, n = n0 + r;
        Qs[i] = n < N ? qb[(size_t)(d * SG_H + h) * N + n] * 0.125f : 0.f;          // / dim ** .5, dim = 64
    }
    float o[4][4], mrow[4], lrow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mrow[i] = -INFINITY; lrow[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    }
    for (int m0 = 0; m0 < M; m0 += BK) {
        __syncthreads();
        for (int i = tid; i < BK * SG_D; i += 256) {
            const int d = i >> 6, c = i & 63, m = m0 + c;
            const bool ok = m < M;
            Ks[i] = ok ? kb[(size_t)(d * SG_H + h) * M + m] : 0.f;
            Vs[c * VP + d] = ok ? vb[(size_t)(d * SG_H + h) * M + m] : 0.f;
        }
        __syncthreads();
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
        for (int d = 0; d < SG_D; ++d) {
            const float4 a = *reinterpret_cast<const float4*>(Qs + d * BQ + ty * 4);
            const float4 bb = *reinterpret_cast<const float4*>(Ks + d * BK + tx * 4);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (m0 + tx * 4 + j >= M) {
#pragma unroll
                for (int i = 0; i < 4; ++i) s[i][j] = -INFINITY;
            }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float rmax = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
#pragma unroll
            for (int w = 8; w > 0; w >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, w));
            const float mnew = fmaxf(mrow[i], rmax);                 // finite: every key tile has at least one valid column
            const float sc = expf(mrow[i] - mnew);
            float p[4], rsum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) { p[j] = expf(s[i][j] - mnew); rsum += p[j]; }
#pragma unroll
            for (int w = 8; w > 0; w >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, w);
            lrow[i] = lrow[i] * sc + rsum;
            mrow[i] = mnew;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= sc;
            *reinterpret_cast<float4*>(Ps + (ty * 4 + i) * VP + tx * 4) = make_float4(p[0], p[1], p[2], p[3]);
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < BK; ++c) {
            const float4 vv = *reinterpret_cast<const float4*>(Vs + c * VP + tx * 4);
            const float vj[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float pi = Ps[(ty * 4 + i) * VP + c];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pi, vj[j], o[i][j]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float inv = 1.f / lrow[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) Qs[(tx * 4 + j) * BQ + ty * 4 + i] = o[i][j] * inv;       // [d][r]
    }
    __syncthreads();
    float* ob = out + (size_t)b * SG_C * N;
    for (int i = tid; i < BQ * SG_D; i += 256) {
        const int d = i >> 6, r = i & 63, n = n0 + r;
        if (n < N) ob[(size_t)(d * SG_H + h) * N + n] = Qs[i];
    }
}

// ---- attention on the tensor cores ----------------------------------------------------------------------------------
// k_sg_attention_tc: one CTA per (128 queries, head, batch element); keys in chunks of 128.  Warp roles (13 warps):
//   warps 0-7   softmax: TWO threads per query row (warp w and w + 4 share TMEM lane quarter w % 4; each owns 64 of the 128
//               key columns of a chunk and 32 of the 64 output dims)
//   warp 8      MMA issue
//   warps 9-12  loaders: transpose chunk c + 1 of K [key][dim] and V [dim][key] into 128-byte-swizzled K-major fp16 (hi, lo)
//               operand slabs (two buffers) while chunk c is processed
// Per chunk: S = Q K^T on tcgen05 (M 128, N 128, K 64) into one of TWO S accumulators in TMEM (S(c+1) is issued while the
// softmax of chunk c runs) -> each softmax thread reads its half row twice (maximum, exchanged with its partner through
// shared memory; then p = exp2(s - max) and its partial row sum), rescales its half of the O row in TMEM when the maximum
// moved, writes P as the next A operand -> O += P V on tcgen05 (M 128, N 64, K 128).
// Every product is the 3-term split (a_hi b_hi + a_lo b_hi + a_hi b_lo, fp32 accumulation): 2^-22 relative, so the match
// decisions downstream see fp32-grade attention.  Q is pre-scaled by log2(e) / 8.
namespace tca {
using namespace oetr::tc;
constexpr int SOFTMAX = 256, WARP_MMA = 8, LOADERS = 128;
constexpr int THREADS = SOFTMAX + 32 + LOADERS;    // 416 (13 warps: 128 registers per thread)
constexpr uint32_t SLAB16 = 128 * 128;             // [128 rows x 64 fp16]
constexpr uint32_t SLAB8 = 64 * 128;               // [64 rows x 64 fp16]
constexpr uint32_t SM_QH = 0, SM_QL = SM_QH + SLAB16;
constexpr uint32_t SM_PH = SM_QL + SLAB16, SM_PL = SM_PH + 2 * SLAB16;
constexpr uint32_t SM_KV = SM_PL + 2 * SLAB16;     // two buffers of {K hi, K lo (16 KB each), V hi, V lo (2 x 8 KB each)} = 64 KB
constexpr uint32_t KV_BUF = 4 * SLAB16, KV_KH = 0, KV_KL = SLAB16, KV_VH = 2 * SLAB16, KV_VL = 3 * SLAB16;
constexpr uint32_t SM_X = SM_KV + 2 * KV_BUF;      // float[2 parities][2 halves][128 rows]: row-maximum exchange; then row sums
constexpr uint32_t SM_BAR = SM_X + 2 * 2 * 128 * 4;
constexpr uint32_t SM_TOTAL = SM_BAR + 128;
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");
struct Bars { uint64_t kv_full[2], kv_free[2], q_full, s[2], p, o; uint32_t tmem, pad; };
static_assert(sizeof(Bars) <= 128, "Bars");
constexpr uint32_t IDESC_S = umma_idesc_f16(128, 128, 0, 0), IDESC_O = umma_idesc_f16(128, 64, 0, 0);
constexpr float Q_SCALE = 0.125f * 1.4426950408889634f;
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

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
// 32 consecutive K-columns (half 0 / 1 of the 64) of row r of a (hi, lo) slab pair
__device__ __forceinline__ void store_row32h(uint8_t* hi, uint8_t* lo, uint32_t r, uint32_t half, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 h, l;
        split8(&v[8 * j], h, l);
        const uint32_t off = slab_chunk_off(r, half * 4 + j);
        *reinterpret_cast<uint4*>(hi + off) = h;
        *reinterpret_cast<uint4*>(lo + off) = l;
    }
}
}  // namespace tca

__global__ void __launch_bounds__(tca::THREADS, 1) k_sg_attention_tc(const float* __restrict__ q, const float* __restrict__ k,
                                                                     const float* __restrict__ v, float* __restrict__ out, int N, int M) {
    using namespace tca;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // key split: the CTAs of a cluster (1, 2 or 4 along x) share a query tile and take consecutive ranges of key chunks;
    // rank 0 merges the partial (max, sum, O) of its peers at the end
    uint32_t crank, csize;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    const int h = blockIdx.y, b = blockIdx.z, n0 = (int)(blockIdx.x / csize) * 128;
    const uint32_t sb = smem_u32(smem);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->kv_full[i], LOADERS); mbar_init(&bars->kv_free[i], 1); mbar_init(&bars->s[i], 1); }
        mbar_init(&bars->q_full, SOFTMAX); mbar_init(&bars->p, SOFTMAX); mbar_init(&bars->o, 1);
        fence_mbar_init();
    }
    if (warp == WARP_MMA) tmem_alloc(&bars->tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem, TS = tmem, TO = tmem + 256;          // S0 | S1 | O
    const int all_chunks = (M + 127) / 128, per = (all_chunks + (int)csize - 1) / (int)csize;
    const int c_begin = (int)crank * per;
    const int chunks = max(0, min(all_chunks, c_begin + per) - c_begin);       // this CTA's chunks (the host keeps csize <= all_chunks)

    if (warp > WARP_MMA) {
        const int row = tid - (SOFTMAX + 32);                                   // 0..127
        const float* kb = k + (size_t)b * SG_C * M;
        const float* vb = v + (size_t)b * SG_C * M;
        const bool vec_ok = (M & 3) == 0 && (reinterpret_cast<uintptr_t>(vb) & 15) == 0;
        for (int c = 0; c < chunks; ++c) {
            const int m0 = (c_begin + c) * 128, bs = c & 1;
            uint8_t* buf = smem + SM_KV + bs * KV_BUF;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {                              // K: key m0 + row, dims half * 32 .. (coalesced over keys)
                float x[32];
                const int key = m0 + row;
#pragma unroll
                for (int d = 0; d < 32; ++d) x[d] = key < M ? __ldg(kb + (size_t)((half * 32 + d) * SG_H + h) * M + key) : 0.f;
                if (half == 0 && c >= 2) mbar_wait(&bars->kv_free[bs], ((c >> 1) - 1) & 1, nullptr);
                store_row32h(buf + KV_KH, buf + KV_KL, row, half, x);
            }
            // V: operand rows = dims, 64 keys (128 B) per slab row.  One item = one 16-byte chunk (8 keys) of a row; the 8
            // lanes of a row read 256 contiguous bytes (two 16-byte loads each when the row is 16-byte aligned)
#pragma unroll 1
            for (int g4 = 0; g4 < 2; ++g4) {                                    // four items per step: 8 independent 16-byte loads in flight
                float x[4][8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int item = (g4 * 4 + u) * LOADERS + row, ci = item & 7, rid = item >> 3, d = rid >> 1, sl = rid & 1;
                    const int key0 = m0 + sl * 64 + ci * 8;
                    const float* src = vb + (size_t)(d * SG_H + h) * M + key0;
                    if (vec_ok && key0 + 8 <= M) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), bq = __ldg(reinterpret_cast<const float4*>(src) + 1);
                        x[u][0] = a.x; x[u][1] = a.y; x[u][2] = a.z; x[u][3] = a.w; x[u][4] = bq.x; x[u][5] = bq.y; x[u][6] = bq.z; x[u][7] = bq.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[u][j] = key0 + j < M ? __ldg(src + j) : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int item = (g4 * 4 + u) * LOADERS + row, ci = item & 7, rid = item >> 3, d = rid >> 1, sl = rid & 1;
                    uint4 hh, ll;
                    split8(x[u], hh, ll);
                    const uint32_t off = sl * SLAB8 + slab_chunk_off(d, ci);
                    *reinterpret_cast<uint4*>(buf + KV_VH + off) = hh;
                    *reinterpret_cast<uint4*>(buf + KV_VL + off) = ll;
                }
            }
            fence_async_smem();
            mbar_arrive(&bars->kv_full[bs]);
        }
    } else if (warp == WARP_MMA) {
        if (lane == 0) {
            auto issue_s = [&](int c) {                                         // S(c) = Q K(c)^T: Qhi.Khi, Qlo.Khi, Qhi.Klo
                const uint32_t kvb = sb + SM_KV + (c & 1) * KV_BUF;
                mbar_wait(&bars->kv_full[c & 1], (c >> 1) & 1, nullptr);
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const uint32_t a = sb + (t == 1 ? SM_QL : SM_QH), bb = kvb + (t == 2 ? KV_KL : KV_KH);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(TS + (c & 1) * 128, umma_desc(a + ks * 32, 16, ATOM_BYTES), umma_desc(bb + ks * 32, 16, ATOM_BYTES),
                                 IDESC_S, (t > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(&bars->s[c & 1]);
            };
            mbar_wait(&bars->q_full, 0, nullptr);
            if (chunks > 0) issue_s(0);
            for (int c = 0; c < chunks; ++c) {
                // S(c+1) goes out while the softmax of chunk c runs: its accumulator was last read for chunk c - 1, whose
                // readers arrived on p(c-1) (waited for below, in the previous iteration)
                if (c + 1 < chunks) issue_s(c + 1);
                const uint32_t kvb = sb + SM_KV + (c & 1) * KV_BUF;
                mbar_wait(&bars->p, c & 1, nullptr);
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < 3; ++t) {                                   // Phi.Vhi, Plo.Vhi, Phi.Vlo
                    const uint32_t a = sb + (t == 1 ? SM_PL : SM_PH), bb = kvb + (t == 2 ? KV_VL : KV_VH);
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_f16(TO, umma_desc(a + sl * SLAB16 + ks * 32, 16, ATOM_BYTES),
                                     umma_desc(bb + sl * SLAB8 + ks * 32, 16, ATOM_BYTES), IDESC_O,
                                     (c > 0 || t > 0 || sl > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(&bars->o);
                umma_commit(&bars->kv_free[c & 1]);
            }
        }
        __syncwarp();
    } else {
        const int quad = warp & 3, hf = warp >> 2;                              // TMEM lane quarter; column half of the row
        const int r = quad * 32 + lane;                                         // query row = TMEM lane
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        float* xch = reinterpret_cast<float*>(smem + SM_X);
        {
            const float* qb = q + (size_t)b * SG_C * N;
            const int n = n0 + r;
            float x[32];
#pragma unroll
            for (int d = 0; d < 32; ++d) x[d] = n < N ? __ldg(qb + (size_t)((hf * 32 + d) * SG_H + h) * N + n) * Q_SCALE : 0.f;
            store_row32h(smem + SM_QH, smem + SM_QL, r, hf, x);
            fence_async_smem();
        }
        mbar_arrive(&bars->q_full);                                             // Q operand published to the MMA warp
        float m_run = -INFINITY, l_part = 0.f;
        for (int c = 0; c < chunks; ++c) {
            const int m0 = (c_begin + c) * 128;
            const uint32_t ts = TS + (c & 1) * 128 + lane_addr + hf * 64;
            mbar_wait(&bars->s[c & 1], (c >> 1) & 1, nullptr);
            tc_fence_after();
            const int valid = (M - m0 < 128 ? M - m0 : 128) - hf * 64;          // valid columns of this thread's half (may be <= 0)
            float rmax = -INFINITY;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                float sv[32];
                tmem_ld32(ts + cc * 32, sv);
#pragma unroll
                for (int j = 0; j < 32; ++j) if (cc * 32 + j < valid) rmax = fmaxf(rmax, sv[j]);
            }
            xch[((c & 1) * 2 + hf) * 128 + r] = rmax;
            named_bar_sync(1 + quad, 64);                                       // the two warps of this lane quarter
            const float m_new = fmaxf(m_run, fmaxf(rmax, xch[((c & 1) * 2 + (hf ^ 1)) * 128 + r]));
            const float sc = ex2(m_run - m_new);                                // 0 on the first chunk
            if (c > 0) { mbar_wait(&bars->o, (c - 1) & 1, nullptr); tc_fence_after(); }   // P and O of the previous chunk are free
            float rsum = 0.f;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                float sv[32];
                tmem_ld32(ts + cc * 32, sv);
#pragma unroll
                for (int j = 0; j < 32; ++j) { sv[j] = cc * 32 + j < valid ? ex2(sv[j] - m_new) : 0.f; rsum += sv[j]; }
                store_row32h(smem + SM_PH + hf * SLAB16, smem + SM_PL + hf * SLAB16, r, cc, sv);
            }
            l_part = l_part * sc + rsum;
            m_run = m_new;
            if (c > 0) {                                                        // rescale this thread's half of the O row in TMEM
                float ov[32];
                tmem_ld32(TO + lane_addr + hf * 32, ov);
#pragma unroll
                for (int j = 0; j < 32; ++j) ov[j] *= sc;
                tmem_st32(TO + lane_addr + hf * 32, ov);
                tmem_st_wait();
            }
            tc_fence_before();
            fence_async_smem();
            mbar_arrive(&bars->p);
        }
        // row sum = the two partial sums (same running maximum in both threads)
        xch[hf * 128 + r] = l_part;
        named_bar_sync(1 + quad, 64);
        float l_tot = l_part + xch[(hf ^ 1) * 128 + r];
        float ov[32];
        if (chunks > 0) {
            mbar_wait(&bars->o, (chunks - 1) & 1, nullptr);
            tc_fence_after();
            tmem_ld32(TO + lane_addr + hf * 32, ov);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) ov[j] = 0.f;
        }
        // merge buffer in rank 0's (now idle) K/V area, per peer p = 1 .. csize-1: M[128] | L[128] | O[128][64]
        float* mb = reinterpret_cast<float*>(smem + SM_KV);
        constexpr int PEER_FLOATS = 128 * 2 + 128 * 64;
        if (csize > 1) {
            tc_fence_before();
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");      // every CTA is done with its own shared memory
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            if (crank != 0) {
                const uint32_t base = smem_u32(mb + (size_t)(crank - 1) * PEER_FLOATS);
                uint32_t ra;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(base), "r"(0u));
                if (hf == 0) {
                    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra + (uint32_t)r * 4u), "f"(m_run) : "memory");
                    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra + (uint32_t)(128 + r) * 4u), "f"(l_tot) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ra + (uint32_t)(256 + r * 64 + hf * 32 + j) * 4u),
                                 "f"(ov[j]), "f"(ov[j + 1]), "f"(ov[j + 2]), "f"(ov[j + 3]) : "memory");
            }
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
        if (crank == 0) {
            float m_all = m_run;
            for (uint32_t pe = 1; pe < csize; ++pe) m_all = fmaxf(m_all, mb[(pe - 1) * PEER_FLOATS + r]);
            const float s0 = ex2(m_run - m_all);
            l_tot *= s0;
#pragma unroll
            for (int j = 0; j < 32; ++j) ov[j] *= s0;
            for (uint32_t pe = 1; pe < csize; ++pe) {
                const float* pb = mb + (pe - 1) * PEER_FLOATS;
                const float sp = ex2(pb[r] - m_all);
                l_tot = fmaf(pb[128 + r], sp, l_tot);
#pragma unroll
                for (int j = 0; j < 32; ++j) ov[j] = fmaf(pb[256 + r * 64 + hf * 32 + j], sp, ov[j]);
            }
            const float inv = 1.f / l_tot;
            float* ob = out + (size_t)b * SG_C * N;
            const int n = n0 + r;
            if (n < N) {
#pragma unroll
                for (int j = 0; j < 32; ++j) ob[(size_t)((hf * 32 + j) * SG_H + h) * N + n] = ov[j] * inv;
            }
        }
    }
    if (warp >= WARP_MMA && csize > 1) {         // the loader and MMA warps take part in the two cluster barriers
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tmem, 512);
}

// ---- log-space optimal transport ------------------------------------------------------------------------------------
// All Sinkhorn arithmetic runs in the log2 domain (values scaled by log2(e) on the fly): exp is then one MUFU (ex2.approx,
// 2 ulp) and the results are converted back by k_sg_finish; mathematically identical to the reference's natural-log form.
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
struct LSE { float m, s; };
__device__ __forceinline__ void lse_merge(LSE& a, const LSE& b) {
    const float m = fmaxf(a.m, b.m);
    if (m == -INFINITY) return;
    a.s = a.s * ex2a(a.m - m) + b.s * ex2a(b.m - m);
    a.m = m;
}
// one chunk of 8 independent elements: local maximum first, then 8 independent exponentials (no serial rescale chain)
__device__ __forceinline__ void lse_chunk(LSE& a, const float (&x)[8]) {
    float cm = x[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) cm = fmaxf(cm, x[i]);
    if (cm == -INFINITY) return;
    const float m = fmaxf(a.m, cm);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += ex2a(x[i] - m);
    a.s = a.s * ex2a(a.m - m) + sum;
    a.m = m;
}
// all CTAs of a cooperative launch: monotonic ticket barrier on a zero-initialised counter
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}
// One half-iteration over the rows of the augmented matrix whose inner part `mat` is [R][Cn] row-major (the scores for the
// u-update, their transpose for the v-update) and whose dustbin row / column is the scalar alpha:
//     out[row] = log2_marg[row] - log2 sum_c 2^(Z[row][c] log2e + in[c])          (in / out in the log2 domain)
// A block handles two rows at a time, four warps per row (a quarter of the columns each, 8 independent coalesced loads in
// flight per lane), merged through shared memory.  `in` was written by other CTAs of the same launch: read through L2 (ld.cg).
__device__ __forceinline__ void sg_half(const float* __restrict__ mat, float alpha2, const float* in, float* out, int batch, int R,
                                        int Cn, float marg_inner, float marg_bin, int first, LSE (*sh)[4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = warp >> 2, quarter = warp & 3;
    const int rows = batch * (R + 1);
    for (int base = blockIdx.x * 2; base < rows; base += gridDim.x * 2) {
        const int r = base + grp;
        LSE a{-INFINITY, 0.f};
        bool inner = false;
        if (r < rows) {
            const int b = r / (R + 1), row = r - b * (R + 1);
            inner = row < R;
            const float* mr = mat + (size_t)b * R * Cn + (size_t)row * Cn;
            const float* iv = in + (size_t)b * (Cn + 1);
            for (int c0 = quarter * 256; c0 <= Cn; c0 += 1024) {
                float x[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int c = c0 + t * 32 + lane;
                    float z = -INFINITY;
                    if (c <= Cn) z = ((inner && c < Cn) ? __ldg(mr + c) * LOG2E : alpha2) + (first ? 0.f : __ldcg(iv + c));
                    x[t] = z;
                }
                lse_chunk(a, x);
            }
#pragma unroll
            for (int w = 16; w > 0; w >>= 1) {
                LSE o{__shfl_xor_sync(0xffffffffu, a.m, w), __shfl_xor_sync(0xffffffffu, a.s, w)};
                lse_merge(a, o);
            }
            if (lane == 0) sh[grp][quarter] = a;
        }
        __syncthreads();
        if (r < rows && quarter == 0 && lane == 0) {
#pragma unroll
            for (int t = 1; t < 4; ++t) lse_merge(a, sh[grp][t]);
            out[r] = (inner ? marg_inner : marg_bin) - (a.m + log2f(a.s));
        }
        __syncthreads();
    }
}
// ALL Sinkhorn iterations in one persistent cooperative kernel: a grid barrier separates the half-iterations (one launch
// instead of 2 x iters: at <= 1024 keypoints the launches themselves were the cost)
__global__ void __launch_bounds__(256) k_sg_sinkhorn(const float* __restrict__ scores, const float* __restrict__ scores_t, float alpha,
                                                     float* u, float* v, int batch, int m, int n, float norm, int iters,
                                                     unsigned int* counter) {
    __shared__ LSE sh[2][4];
    const float norm2 = norm * LOG2E, alpha2 = alpha * LOG2E;
    const float bin_u = log2f((float)n) + norm2, bin_v = log2f((float)m) + norm2;
    unsigned int phase = 0;
    for (int it = 0; it < iters; ++it) {
        sg_half(scores, alpha2, v, u, batch, m, n, norm2, bin_u, it == 0, sh);
        grid_barrier(counter, ++phase * gridDim.x);
        sg_half(scores_t, alpha2, u, v, batch, n, m, norm2, bin_v, 0, sh);
        grid_barrier(counter, ++phase * gridDim.x);
    }
}
// scores [m][n] -> transposed copy [n][m] (32 x 32 tiles through shared memory), once per call
__global__ void __launch_bounds__(256) k_sg_transpose(const float* __restrict__ src, float* __restrict__ dst, int m, int n) {
    __shared__ float t[32][33];
    const int b = blockIdx.z, j0 = blockIdx.x * 32, i0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* s = src + (size_t)b * m * n;
    float* d = dst + (size_t)b * m * n;
    for (int r = ty; r < 32; r += 8)
        if (i0 + r < m && j0 + tx < n) t[r][tx] = s[(size_t)(i0 + r) * n + j0 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
        if (j0 + r < n && i0 + tx < m) d[(size_t)(j0 + r) * m + i0 + tx] = t[tx][r];
}
// out[i][j] = Z[i][j] + (u2[i] + v2[j]) ln 2 - norm        ([m+1][n+1], the matrix the reference returns; u2, v2: log2 domain)
__global__ void __launch_bounds__(256) k_sg_finish(const float* __restrict__ scores, float alpha, const float* __restrict__ uu,
                                                   const float* __restrict__ vv, float* __restrict__ out, int m, int n, float norm) {
    const int b = blockIdx.z, j = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
    if (j > n) return;
    const float z = (i < m && j < n) ? scores[(size_t)b * m * n + (size_t)i * n + j] : alpha;
    out[(size_t)b * (m + 1) * (n + 1) + (size_t)i * (n + 1) + j] = z + (uu[(size_t)b * (m + 1) + i] + vv[(size_t)b * (n + 1) + j]) * LN2 - norm;
}

thread_local char g_serr[256] = "";
int sfail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_serr, sizeof(g_serr), fmt, ap);
    va_end(ap);
    return code;
}
std::mutex g_sg_mu;
bool g_sg_attr[64] = {};

}  // namespace

extern "C" {

const char* oetr_sg_last_error(void) { return g_serr; }

int oetr_sg_attention(const float* query, const float* key, const float* value, float* out, int batch, int n, int m, int mode,
                      void* stream) {
    if (!query || !key || !value || !out) return sfail(OETR_E_ARG, "oetr_sg_attention: null argument");
    if (batch < 1 || n < 1 || m < 1 || batch > 65535) return sfail(OETR_E_SHAPE, "oetr_sg_attention: batch %d, %d queries, %d keys", batch, n, m);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_sg_mu);
        if (dev < 64 && !g_sg_attr[dev]) {
            e = cudaFuncSetAttribute(k_sg_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sg_attention_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tca::SM_TOTAL);
            if (e == cudaSuccess) g_sg_attr[dev] = true;
        }
    }
    if (e != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_attention: %s", cudaGetErrorString(e));
    if (mode != OETR_SG_TENSOR && mode != OETR_SG_FP32) return sfail(OETR_E_ARG, "oetr_sg_attention: mode %d", mode);
    if (mode == OETR_SG_TENSOR) {
        // key split over a cluster of 1, 2 or 4 CTAs so that small problems still fill the GPU (2048 keypoints: 64 query
        // tile x head CTAs on 148 SMs)
        const int qt = (n + 127) / 128, chunks = (m + 127) / 128, base = qt * SG_H * batch;
        int split = 1;
        while (split < 4 && base * split * 2 <= 160 && split * 2 <= chunks) split *= 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(qt * split, SG_H, batch); cfg.blockDim = dim3(tca::THREADS); cfg.dynamicSmemBytes = tca::SM_TOTAL;
        cfg.stream = static_cast<cudaStream_t>(stream);
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = split; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, k_sg_attention_tc, query, key, value, out, n, m);
        if (e != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_attention: %s", cudaGetErrorString(e));
    } else
        k_sg_attention<<<dim3((n + BQ - 1) / BQ, SG_H, batch), 256, ATT_SMEM, static_cast<cudaStream_t>(stream)>>>(query, key, value, out, n, m);
    e = cudaGetLastError();
    if (e != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_attention: %s", cudaGetErrorString(e));
    return OETR_OK;
}

// barrier counter (256 B) | u [batch][m+1] | v [batch][n+1] | transposed scores [batch][n][m]
size_t oetr_sg_transport_workspace_bytes(int batch, int m, int n) {
    if (batch < 1 || m < 1 || n < 1) return 0;
    return 256 + (size_t)batch * ((size_t)(m + n + 2) + (size_t)m * n) * sizeof(float);
}

int oetr_sg_optimal_transport(const float* scores, float alpha, int iters, float* out, int batch, int m, int n, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (!scores || !out || !workspace) return sfail(OETR_E_ARG, "oetr_sg_optimal_transport: null argument");
    if (batch < 1 || m < 1 || n < 1 || iters < 0 || batch > 65535 || m > 65534)
        return sfail(OETR_E_SHAPE, "oetr_sg_optimal_transport: batch %d, %d x %d scores, %d iterations", batch, m, n, iters);
    if (workspace_bytes < oetr_sg_transport_workspace_bytes(batch, m, n))
        return sfail(OETR_E_NOMEM, "oetr_sg_optimal_transport: workspace of %zu bytes, %zu needed", workspace_bytes,
                     oetr_sg_transport_workspace_bytes(batch, m, n));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned int* counter = static_cast<unsigned int*>(workspace);
    float* u = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
    float* v = u + (size_t)batch * (m + 1);
    float* st = v + (size_t)batch * (n + 1);
    const float norm = -logf((float)(m + n));
    cudaError_t e0 = cudaMemsetAsync(workspace, 0, 256 + (iters == 0 ? (size_t)batch * (m + n + 2) * sizeof(float) : 0), s);
    if (e0 != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_optimal_transport: %s", cudaGetErrorString(e0));
    if (iters > 0) {
        k_sg_transpose<<<dim3((n + 31) / 32, (m + 31) / 32, batch), 256, 0, s>>>(scores, st, m, n);
        int dev = 0, sms = 0, per_sm = 0;
        e0 = cudaGetDevice(&dev);
        if (e0 == cudaSuccess) e0 = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e0 == cudaSuccess) e0 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sg_sinkhorn, 256, 0);
        if (e0 != cudaSuccess || per_sm < 1) return sfail(OETR_E_CUDA, "oetr_sg_optimal_transport: occupancy query failed");
        const long rows = (long)batch * ((m > n ? m : n) + 1);
        long grid = (rows + 1) / 2;                              // two rows per block and step
        // as many co-resident blocks as fit: a sweep of 1 / 2 / 4 / 8 blocks per SM gave 6.6 / 4.1 / 2.7 / 2.5 ms for 100
        // iterations at 2048 keypoints (the rows are latency-bound, the grid barrier costs ~4 us whatever the block count)
        if (grid > (long)sms * per_sm) grid = (long)sms * per_sm;
        const float* sc_ = scores; const float* st_ = st;
        void* kargs[] = {(void*)&sc_, (void*)&st_, (void*)&alpha, (void*)&u, (void*)&v, (void*)&batch, (void*)&m, (void*)&n,
                         (void*)&norm, (void*)&iters, (void*)&counter};
        e0 = cudaLaunchCooperativeKernel((const void*)k_sg_sinkhorn, dim3((unsigned)grid), dim3(256), kargs, 0, s);
        if (e0 != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_optimal_transport: cooperative launch: %s", cudaGetErrorString(e0));
    }
    k_sg_finish<<<dim3((n + 1 + 255) / 256, m + 1, batch), 256, 0, s>>>(scores, alpha, u, v, out, m, n, norm);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_optimal_transport: %s", cudaGetErrorString(e));
    return OETR_OK;
}

}  // extern "C"
