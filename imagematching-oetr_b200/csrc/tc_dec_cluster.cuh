// k_decoder_cl: the 2-layer query decoder + size regression (transformer.py:224-284,361-381; src/model.py:188-191) on
// thread-block CLUSTERS.  The one-CTA version (k_decoder, tc_head.cuh) streams all 5.8 MB of transposed fp32 weights
// through ONE SM per two query tokens: 100 us of pure weight streaming on 8 SMs per sub-batch, 10 % of a sub-batch's launch
// chain.  Here a cluster of 8 CTAs serves 16 query tokens: CTA `rank` owns output columns [32 rank, 32 rank + 32) of every
// projection -- exactly attention head `rank` -- so it streams 1/8 of every weight matrix (a pre-sliced, contiguous
// per-rank stream: 21 bulk copies of 32 KB), and the per-head arithmetic of both attentions is local to a CTA (and to a
// warp: lanes = the head's 32 channels).  Only the inputs of the next projection are exchanged: each CTA writes its output
// slice to its own shared memory, the cluster synchronises (barrier.cluster, 13 times per forward) and every CTA pulls the
// 8 slices over DSMEM (ld.shared::cluster, 16 bytes per load).  LayerNorms run redundantly in every CTA on the gathered rows.
// Part of the tcgen05 path's head; compiled into tc_kernels.cu.
#pragma once
#include "tc_head.cuh"

namespace oetr {

constexpr int DCL_RANKS = 8, DCL_ROWS = 16, DCL_THREADS = 256, DCL_STAGES = 3;
constexpr uint32_t DCL_CHUNK = 32768;
constexpr int DCL_CHUNK_FLOATS = DCL_CHUNK / 4;
constexpr size_t DCL_LAYER_FLOATS = (size_t)6 * C * 32 + (size_t)C * 64 + (size_t)FF * 32;       // per rank and layer
constexpr size_t DCL_RANK_FLOATS = N_DEC * DCL_LAYER_FLOATS + (size_t)C * 32;                    // + tlbr_reg.0 slice
constexpr uint32_t DCL_TOTAL_CHUNKS = (uint32_t)(DCL_RANK_FLOATS * 4 / DCL_CHUNK);               // 21
static_assert(DCL_RANK_FLOATS * 4 % DCL_CHUNK == 0, "whole chunks");
// shared memory: ring | V[16][512] the gathered vector | S[2][16][64] this CTA's own output slice (what peers pull) |
// u[16][256] | a[16][256] | barriers
constexpr uint32_t DCL_SM_V = DCL_STAGES * DCL_CHUNK;
constexpr uint32_t DCL_SM_S = DCL_SM_V + DCL_ROWS * FF * 4;
constexpr uint32_t DCL_SM_U = DCL_SM_S + 2 * DCL_ROWS * 64 * 4;
constexpr uint32_t DCL_SM_A = DCL_SM_U + DCL_ROWS * C * 4;
// staged small vectors per layer: ln1_g ln1_b ln2_g ln2_b ln3_g ln3_b sa_bq sa_bk sa_bv ca_bq (10 x 256), then qe1 | qe2
constexpr int DCL_PV = 10 * C;
constexpr uint32_t DCL_SM_P = DCL_SM_A + DCL_ROWS * C * 4;
constexpr uint32_t DCL_SM_BAR = DCL_SM_P + (N_DEC * DCL_PV + 2 * C) * 4;
constexpr uint32_t DCL_SMEM = DCL_SM_BAR + 64;

struct DecClParams {
    DecLayerT layer[N_DEC];
    const float* qe;            // query_embed1 | query_embed2
    const float* kvs;           // [N_DEC][2B][KVS]
    float* hs;                  // out [2B][256]
    const float* wts;           // [8 ranks][DCL_RANK_FLOATS] sliced transposed weights in consumption order
    const float *tl_w2, *tl_b2;
    float* tlbr;                // out [2B][4]
    int B;
};

// W[N][K] row-major -> per-rank slices WTs[rank][k][nl] = W[rank * NL + nl][k]   (NL = N / 8); dst = slice of rank 0,
// rank_stride floats between ranks.  k rows [k0, k0 + kn) only (a chunk of the stream).
__global__ void k_slice_t(const float* __restrict__ W, int K, int NL, int k0, int kn, float* __restrict__ dst, size_t rank_stride) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= DCL_RANKS * kn * NL) return;
    const int nl = idx % NL, kk = (idx / NL) % kn, rank = idx / (NL * kn);
    dst[(size_t)rank * rank_stride + (size_t)kk * NL + nl] = W[(size_t)(rank * NL + nl) * K + k0 + kk];
}

struct DclRing { uint64_t* full; uint64_t* empty; const float* stage0; const uint8_t* src; uint32_t cons, prod; };

__device__ __forceinline__ void dcl_refill(DclRing& r, uint8_t* smem) {           // thread 0 only
    while (r.prod < DCL_TOTAL_CHUNKS && r.prod < r.cons + DCL_STAGES) {
        const int st = r.prod % DCL_STAGES;
        if (r.prod >= DCL_STAGES) mbar_wait(&r.empty[st], ((r.prod / DCL_STAGES) - 1) & 1, nullptr);
        mbar_arrive_expect_tx(&r.full[st], DCL_CHUNK);
        bulk_g2s(smem + (size_t)st * DCL_CHUNK, r.src + (size_t)r.prod * DCL_CHUNK, DCL_CHUNK, &r.full[st]);
        ++r.prod;
    }
}
// acc[i][c] += sum_k ws[k][lane + 32 c] * xin[row_i][k] for this thread's two rows; the weight slice comes from the ring
template <int K, int NL>
__device__ __forceinline__ void dcl_matvec(DclRing& ring, uint8_t* smem, const float* x0, const float* x1, float (&acc)[2][NL / 32]) {
    constexpr int ROWS = DCL_CHUNK_FLOATS / NL;           // k rows per chunk
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < NL / 32; ++c) acc[i][c] = 0.f;
    for (int k0 = 0; k0 < K; k0 += ROWS) {
        if (threadIdx.x == 0) dcl_refill(ring, smem);
        const int st = ring.cons % DCL_STAGES;
        mbar_wait(&ring.full[st], (ring.cons / DCL_STAGES) & 1, nullptr);
        const float* ws = ring.stage0 + (size_t)st * DCL_CHUNK_FLOATS;
#pragma unroll 4
        for (int kk = 0; kk < ROWS; kk += 4) {
            const float4 xa = *reinterpret_cast<const float4*>(x0 + k0 + kk), xb = *reinterpret_cast<const float4*>(x1 + k0 + kk);
#pragma unroll
            for (int c = 0; c < NL / 32; ++c) {
                const float w0 = ws[(kk + 0) * NL + lane + 32 * c], w1 = ws[(kk + 1) * NL + lane + 32 * c];
                const float w2 = ws[(kk + 2) * NL + lane + 32 * c], w3 = ws[(kk + 3) * NL + lane + 32 * c];
                acc[0][c] = fmaf(w3, xa.w, fmaf(w2, xa.z, fmaf(w1, xa.y, fmaf(w0, xa.x, acc[0][c]))));
                acc[1][c] = fmaf(w3, xb.w, fmaf(w2, xb.z, fmaf(w1, xb.y, fmaf(w0, xb.x, acc[1][c]))));
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[st]);
        ++ring.cons;
    }
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// All-gather by PULLING: every CTA has written its own output slice S[par][16][W] (W = 32 or 64 columns) to its own shared
// memory and the cluster has synchronised; each CTA now copies the 8 slices into its local V[16][8 W] with 16-byte
// ld.shared::cluster loads (scalar st.shared::cluster pushes -- 16 per thread -- made the first version no faster than the
// one-CTA kernel).  S is double-buffered: a peer may still be pulling S[par] while this CTA already writes S[par ^ 1].
template <int W>
__device__ __forceinline__ void dcl_pull(float* V, const float* S_par) {
    constexpr int F4_ROW = W / 4, ITEMS = DCL_RANKS * DCL_ROWS * F4_ROW, PER = ITEMS / DCL_THREADS;     // float4 items: (rank, row, quad)
    static_assert(ITEMS % DCL_THREADS == 0, "whole items per thread");
    const uint32_t la = smem_u32(S_par);
    float4 v[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {                                  // all remote loads in flight together
        const int it = i * DCL_THREADS + threadIdx.x;
        const int qd = it % F4_ROW, row = (it / F4_ROW) % DCL_ROWS, rk = it / (F4_ROW * DCL_ROWS);
        uint32_t ra;
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la + (uint32_t)(row * 64 + qd * 4) * 4u), "r"((uint32_t)rk));
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(ra));
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int it = i * DCL_THREADS + threadIdx.x;
        const int qd = it % F4_ROW, row = (it / F4_ROW) % DCL_ROWS, rk = it / (F4_ROW * DCL_ROWS);
        *reinterpret_cast<float4*>(V + row * FF + rk * W + qd * 4) = v[i];
    }
    __syncthreads();
}
// u[row] = LN(src[row]) and a[row] = u[row] + qe(row) for this warp's two rows (src rows have stride FF)
__device__ __forceinline__ void dcl_ln(const float* src, const float* __restrict__ g, const float* __restrict__ b, float* u, float* a,
                                       const float* qe0, const float* qe1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = 2 * warp + i;
        const float* qe = i ? qe1 : qe0;
        float v[8], s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] = src[row * FF + lane + 32 * j]; s += v[j]; }
        const float mu = warp_sum_f(s) * (1.f / C);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[j] - mu; sq = fmaf(d, d, sq); }
        const float rstd = rsqrtf(warp_sum_f(sq) * (1.f / C) + LN_EPS);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            const float y = (v[j] - mu) * rstd * g[c] + b[c];
            u[row * C + c] = y;
            a[row * C + c] = y + qe[c];
        }
    }
}

__global__ void __cluster_dims__(DCL_RANKS, 1, 1) __launch_bounds__(DCL_THREADS, 1) k_decoder_cl(const DecClParams p) {
    extern __shared__ __align__(1024) uint8_t dsm[];
    float* V = reinterpret_cast<float*>(dsm + DCL_SM_V);             // [16][512]: the gathered vector (every CTA has all of it)
    float* S = reinterpret_cast<float*>(dsm + DCL_SM_S);             // [2][16][64]: this CTA's slice of the next gathered vector
    float* u = reinterpret_cast<float*>(dsm + DCL_SM_U);
    float* a = reinterpret_cast<float*>(dsm + DCL_SM_A);
    uint64_t* bars = reinterpret_cast<uint64_t*>(dsm + DCL_SM_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_rank(), group = blockIdx.x / DCL_RANKS;
    const int rows = 2 * p.B, row_base = group * DCL_ROWS;
    const int n = rank * 32 + lane;                                  // this thread's output channel = head `rank`, dim `lane`
    if (tid == 0) {
        for (int i = 0; i < DCL_STAGES; ++i) { mbar_init(&bars[i], 1); mbar_init(&bars[DCL_STAGES + i], 8); }
        fence_mbar_init();
    }
    for (int i = tid; i < DCL_ROWS * FF; i += DCL_THREADS) V[i] = 0.f;            // tgt = zeros (transformer.py:361)
    // every small vector the layers read (LayerNorm parameters, projection biases, query embeddings) is staged in shared
    // memory once, coalesced, under the first weight copies: the layer code then has no dependent global-load latencies
    // (the decoder is a serial chain of 17 small products: those ~1 us stalls were most of the one-CTA kernel's 100 us)
    float* P = reinterpret_cast<float*>(dsm + DCL_SM_P);
    for (int j = 0; j < N_DEC; ++j) {
        const DecLayerT& w = p.layer[j];
        const float* src[10] = {w.ln1_g, w.ln1_b, w.ln2_g, w.ln2_b, w.ln3_g, w.ln3_b, w.sa_bq, w.sa_bk, w.sa_bv, w.ca_bq};
#pragma unroll
        for (int v = 0; v < 10; ++v) P[j * DCL_PV + v * C + tid] = src[v][tid];
    }
    P[N_DEC * DCL_PV + tid] = p.qe[tid];
    P[N_DEC * DCL_PV + C + tid] = p.qe[C + tid];
    __syncthreads();
    DclRing ring{bars, bars + DCL_STAGES, reinterpret_cast<const float*>(dsm),
                 reinterpret_cast<const uint8_t*>(p.wts + (size_t)rank * DCL_RANK_FLOATS), 0u, 0u};
    const int r0 = 2 * warp, r1 = r0 + 1;
    const int g0 = min(row_base + r0, rows - 1), g1 = min(row_base + r1, rows - 1);      // clamped global rows (reads)
    const float* qe0 = P + N_DEC * DCL_PV + (g0 >= p.B ? C : 0);
    const float* qe1 = P + N_DEC * DCL_PV + (g1 >= p.B ? C : 0);
    float t_my[2] = {0.f, 0.f};
    int par = 0;
    // this thread's values of rows r0, r1 (output channel n) -> own slice -> cluster barrier -> every CTA pulls all slices
    // into its V.  All local reads of V (the previous gathered vector) are complete before the pull overwrites it: the
    // values being gathered were computed from it.
    auto gather2 = [&](float v0, float v1) {
        par ^= 1;
        float* Sp = S + par * DCL_ROWS * 64;
        Sp[r0 * 64 + lane] = v0;
        Sp[r1 * 64 + lane] = v1;
        cluster_sync_all();
        dcl_pull<32>(V, Sp);
    };
    auto Vcur = [&]() { return V; };
    for (int j = 0; j < N_DEC; ++j) {
        const float* Pj = P + j * DCL_PV;
        float acc[2][1], kk[2][1], vv[2][1], h2[2][2];
        // the cross-attention summaries of this thread's head and rows: in flight during the self-attention
        float kvr[2][HD], ksr[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float* kv = p.kvs + ((size_t)j * rows + (i ? g1 : g0)) * KVS;
            ksr[i] = __ldg(kv + NH * HD * HD + n);
#pragma unroll
            for (int d = 0; d < HD; ++d) kvr[i][d] = __ldg(kv + (rank * HD + d) * HD + lane);
        }
        // ---- self-attention over the single query token (transformer.py:236-241, linear_attention.py:22-50)
        dcl_ln(Vcur(), Pj + 0 * C, Pj + 1 * C, u, a, qe0, qe1);
        __syncthreads();
        dcl_matvec<C, 32>(ring, dsm, a + r0 * C, a + r1 * C, acc);
        dcl_matvec<C, 32>(ring, dsm, a + r0 * C, a + r1 * C, kk);
        dcl_matvec<C, 32>(ring, dsm, u + r0 * C, u + r1 * C, vv);
        float o2[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float qf = elu1(acc[i][0] + Pj[6 * C + n]), kf = elu1(kk[i][0] + Pj[7 * C + n]);
            const float sden = warp_sum_f(qf * kf);                                   // warp lanes = the head's channels
            o2[i] = (vv[i][0] + Pj[8 * C + n]) * sden / (sden + ATTN_EPS);
        }
        gather2(o2[0], o2[1]);
        dcl_matvec<C, 32>(ring, dsm, Vcur() + r0 * FF, Vcur() + r1 * FF, acc);
        t_my[0] += acc[0][0]; t_my[1] += acc[1][0];
        gather2(t_my[0], t_my[1]);
        // ---- cross-attention into the memory summaries (transformer.py:243-250)
        dcl_ln(Vcur(), Pj + 2 * C, Pj + 3 * C, u, a, qe0, qe1);
        __syncthreads();
        dcl_matvec<C, 32>(ring, dsm, a + r0 * C, a + r1 * C, acc);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float qv = elu1(acc[i][0] + Pj[9 * C + n]);
            const float den = warp_sum_f(qv * ksr[i]);
            float o = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) o = fmaf(__shfl_sync(0xffffffffu, qv, d), kvr[i][d], o);
            o2[i] = o / (den + ATTN_EPS);
        }
        gather2(o2[0], o2[1]);
        dcl_matvec<C, 32>(ring, dsm, Vcur() + r0 * FF, Vcur() + r1 * FF, acc);
        t_my[0] += acc[0][0]; t_my[1] += acc[1][0];
        gather2(t_my[0], t_my[1]);
        // ---- feed-forward (transformer.py:252-254): hidden columns [64 rank, 64 rank + 64) here
        dcl_ln(Vcur(), Pj + 4 * C, Pj + 5 * C, u, a, qe0, qe1);
        __syncthreads();
        dcl_matvec<C, 64>(ring, dsm, u + r0 * C, u + r1 * C, h2);
        {
            par ^= 1;
            float* Sp = S + par * DCL_ROWS * 64;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                Sp[r0 * 64 + lane + 32 * c] = fmaxf(h2[0][c], 0.f);
                Sp[r1 * 64 + lane + 32 * c] = fmaxf(h2[1][c], 0.f);
            }
            cluster_sync_all();
            dcl_pull<64>(V, Sp);
        }
        dcl_matvec<FF, 32>(ring, dsm, Vcur() + r0 * FF, Vcur() + r1 * FF, acc);
        t_my[0] += acc[0][0]; t_my[1] += acc[1][0];
        gather2(t_my[0], t_my[1]);
    }
    if (row_base + r0 < rows) p.hs[(size_t)(row_base + r0) * C + n] = t_my[0];
    if (row_base + r1 < rows) p.hs[(size_t)(row_base + r1) * C + n] = t_my[1];
    // ---- size regression (src/model.py:188-191): sigmoid(W_b relu(W_a hs) + b); output `rank` (< 4) per CTA
    {
        float acc[2][1];
        dcl_matvec<C, 32>(ring, dsm, Vcur() + r0 * FF, Vcur() + r1 * FF, acc);
        gather2(fmaxf(acc[0][0], 0.f), fmaxf(acc[1][0], 0.f));
        if (rank < 4) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float* x = Vcur() + (i ? r1 : r0) * FF;
                float o = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) o = fmaf(p.tl_w2[(size_t)rank * C + lane + 32 * q], x[lane + 32 * q], o);
                o = warp_sum_f(o);
                const int row = row_base + (i ? r1 : r0);
                if (lane == 0 && row < rows) p.tlbr[(size_t)row * 4 + rank] = 1.f / (1.f + expf(-(o + p.tl_b2[rank])));
            }
        }
    }
    cluster_sync_all();                                              // no CTA exits while a peer may still pull from its shared memory
}

}  // namespace oetr
