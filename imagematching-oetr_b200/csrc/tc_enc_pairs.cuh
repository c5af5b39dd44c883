// k_enc2: the experimental CTA-pair (tcgen05 cta_group::2) variant of k_enc, OETR_ENC=2.
// Part of the tcgen05 (OETR_PREC_FP16) path; compiled into tc_kernels.cu (one translation unit: kernels are
// launched from the host code there).
#pragma once
#include "tc_enc.cuh"

namespace oetr {
using namespace tc;

// ---------------------------------------------------------------------------------------------------------
// k_enc2: the same layer kernel on CTA PAIRS (tcgen05 cta_group::2), two CTAs resident per SM.
//
// k_enc is a serial chain per tile (GEMM -> row-wise epilogue -> GEMM ...): the tensor core idles while the row
// warps work and vice versa (tensor pipe 36 % busy), and one CTA owns the whole SM (226 KB of shared memory, all
// of TMEM).  Here a 128-token tile belongs to a cluster of two CTAs on the two SMs of a TPC: each CTA holds 64
// token rows (operand image 64 KB instead of 128) and streams only ITS half of every weight tile (the N rows
// [128*rank, +128); ring 3 x 16 KB instead of 96 KB), and the leader's MMA lane issues M=128 N=256
// tcgen05.mma.cta_group::2 instructions for the pair (64 cycles each: both tensor cores at full rate).  The
// accumulators use the "2x2" TMEM layout (lanes 0-63: columns n < 128, lanes 64-127: n >= 128 of the CTA's 64 rows),
// 128 TMEM columns each, so S0 | S1 need 256 of the 512 columns.  A CTA then needs 113 KB / 256 columns / 320
// threads, TWO CTAs (of different pairs) fit one SM, and while one waits for its epilogue the other one's MMAs run.
//   warps 0-7   row warps: thread <-> (token row 32*(w&1)+lane, columns [128*((w>>1)&1) + 64*(w>>2), +64)), i.e. TMEM
//               lane quarter w%4; the fp32 residual stream stays in registers (64 per thread) as in k_enc
//   warp 8      weight producer (both CTAs, own half), TMEM allocation
//   warp 9      leader: MMA issue for the pair;  peer: relays "my half of the stage has landed" to the leader
// Cross-CTA signalling: row warps of both CTAs arrive (one lane per warp, release.cluster) on the leader's a_full;
// tcgen05.commit multicasts the ring `empty` and accumulator `s_full` arrivals to both CTAs.
// KV = Kf^T V needs tokens as the K dimension, and the pair's tokens are split across CTAs, so the pair MMA is
// used with N split as (own V | peer's V): CTA r keeps the half of the result built from its own V (n-half r)
// and ignores the cross term; the per-CTA partial summaries are added by k_fold / k_sum_partials.
// ---------------------------------------------------------------------------------------------------------
constexpr int HT = 64;                                  // token rows per CTA
constexpr int E2_ROW_THREADS = 256, E2_WARP_PRODUCER = 8, E2_WARP_MMA = 9, E2_THREADS = 320;
constexpr int E2_RING = 3;                              // ring units of ONE 16 KB stage (this CTA's half of a B tile)
constexpr uint32_t E2_SLAB = HT * 128;                  // [64 rows x 64 K] fp16 = 8 KB
constexpr uint32_t E2_IMG = 4 * E2_SLAB;                // 32 KB
constexpr uint32_t E2_AHI = 0, E2_ALO = E2_IMG, E2_RINGOFF = 2 * E2_IMG;
constexpr uint32_t E2_X = E2_RINGOFF + E2_RING * STAGE_BYTES;       // 192-float scratch (LayerNorm / Ksum exchange)
constexpr uint32_t E2_BAR = E2_X + 192 * 4;
constexpr uint32_t E2_TOTAL = E2_BAR + 128;
static_assert(2 * (E2_TOTAL + 1024) <= 228 * 1024, "two CTAs per SM");
constexpr uint32_t IDESC2_N256 = umma_idesc_f16(128, 256, 0, 0);    // pair MMA: M = 128 (64 rows per CTA)
constexpr uint32_t IDESC2_KV = umma_idesc_f16(128, 128, 1, 1);

struct Bars2 {
    uint64_t full[E2_RING], empty[E2_RING];
    uint64_t pfull[E2_RING];   // leader: the peer's half of the stage has landed (relayed by the peer's warp 9)
    uint64_t a_full;           // leader: operand image written by the 16 row warps of the pair
    uint64_t kv_a, kv_b;       // leader: KV round 0 (all 16 warps) / round 1 (the 8 warps of N half 1)
    uint64_t s_full[2];
    uint32_t tmem_base;
    uint32_t pad;
};
static_assert(sizeof(Bars2) <= 128, "Bars2 must fit its reservation");

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier of any CTA of the cluster (address from mapa), release at cluster scope
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// bounded wait, acquire at cluster scope, with a suspend-time hint: a waiting warp sleeps instead of polling, so it
// does not take issue slots from the other CTA resident on the SM
__device__ __forceinline__ void mbar_wait2(uint64_t* bar, uint32_t parity, int* timeout_flag) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 16); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity), "r"(20000u)
            : "memory");
        if (done) return;
        if ((spin & 255) == 255 && timeout_flag && *reinterpret_cast<volatile int*>(timeout_flag)) return;
    }
    if (timeout_flag) atomicExch(timeout_flag, 1);
}
// the same without the hint, for the single producer / relay / MMA lanes (their wake-up latency is on the critical
// path of the weight ring and one polling lane costs nothing)
__device__ __forceinline__ void mbar_wait2_poll(uint64_t* bar, uint32_t parity, int* timeout_flag) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 20); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) return;
        if ((spin & 1023) == 1023 && timeout_flag && *reinterpret_cast<volatile int*>(timeout_flag)) return;
    }
    if (timeout_flag) atomicExch(timeout_flag, 1);
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {     // one warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all MMAs issued so far by this thread -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(E2_THREADS, 2) k_enc2(const EncParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars2* bars = reinterpret_cast<Bars2*>(smem + E2_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int tile = blockIdx.x >> 1;
    const TileInfo ti = tile_info(p.g, tile);
    const uint32_t smem_base = smem_u32(smem);
    const bool dec_mode = p.lnkv_g == nullptr;

    if (tid == 0) {
        for (int i = 0; i < E2_RING; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); mbar_init(&bars->pfull[i], 1); }
        mbar_init(&bars->a_full, 16);
        mbar_init(&bars->kv_a, 16);
        mbar_init(&bars->kv_b, 8);
        mbar_init(&bars->s_full[0], 1);
        mbar_init(&bars->s_full[1], 1);
        fence_mbar_init();
    }
    if (warp == E2_WARP_PRODUCER) tmem_alloc2(&bars->tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // the peer's barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t S0 = tmem, S1 = tmem + 128;

    int src_img = ti.img, src_len = ti.L;
    if (p.cross) { src_img = ti.set == 0 ? p.g.B + ti.b : ti.b; src_len = ti.set == 0 ? p.g.L2 : p.g.L1; }
    // units (16 KB stages) this launch streams per CTA: 8 per GEMM
    const int n_units = (p.do_q ? 6 * 8 : 0) + (p.do_kv ? 2 * 8 : 0);

    if (warp == E2_WARP_PRODUCER) {
        // ------------------------------------------------------------------ weight stream (own N half)
        if (lane == 0) {
#pragma unroll 1
            for (int k = 0; k < 3; ++k)
                for (uint32_t off = blockIdx.x * STAGE_BYTES; off < p.pf_bytes[k]; off += gridDim.x * STAGE_BYTES)
                    bulk_prefetch_l2(static_cast<const uint8_t*>(p.pf_ptr[k]) + off, min(STAGE_BYTES, p.pf_bytes[k] - off));
            uint32_t g = 0;
            auto stream = [&](const __half* src, int ngemms) {
                for (int gm = 0; gm < ngemms; ++gm)
                    for (int ks = 0; ks < 4; ++ks)
                        for (int lo = 0; lo < 2; ++lo, ++g) {
                            const int st = g % E2_RING;
                            mbar_wait2_poll(&bars->empty[st], ((g / E2_RING) & 1) ^ 1, p.flag);
                            mbar_arrive_expect_tx(&bars->full[st], STAGE_BYTES);
                            bulk_g2s(smem + E2_RINGOFF + st * STAGE_BYTES,
                                     src + (size_t)gm * GEMM_HALFS + gemm_stage_off(ks, lo, (int)rank), STAGE_BYTES, &bars->full[st]);
                        }
            };
            if (p.do_q) {
                stream(p.w_q, 1);
                stream(p.mimg + (size_t)src_img * GEMM_HALFS, 1);
                stream(p.w_mlp, 4);
            }
            if (p.do_kv) stream(p.w_kv, 2);
        }
        __syncwarp();
    } else if (warp == E2_WARP_MMA) {
        if (lane == 0 && rank == 1) {
            // ------------------------------------------------------------------ peer: relay "stage landed"
            for (int g = 0; g < n_units; ++g) {
                const int st = g % E2_RING;
                mbar_wait2_poll(&bars->full[st], (g / E2_RING) & 1, p.flag);
                mbar_arrive_cluster(mapa_u32(smem_u32(&bars->pfull[st]), 0));
            }
        } else if (lane == 0) {
            // ------------------------------------------------------------------ leader: MMA issue for the pair
            uint32_t g = 0, na = 0;
            const long long t_begin = clock64();
            long long t_a = 0, t_ring = 0;
            auto wait_unit = [&]() -> uint32_t {
                const int st = g % E2_RING;
                const long long t0 = clock64();
                mbar_wait2_poll(&bars->full[st], (g / E2_RING) & 1, p.flag);
                mbar_wait2_poll(&bars->pfull[st], (g / E2_RING) & 1, p.flag);
                t_ring += clock64() - t0;
                tc_fence_after();
                return smem_base + E2_RINGOFF + st * STAGE_BYTES;
            };
            auto gemm = [&](uint32_t d, bool accumulate, bool wait) {
                if (wait) {
                    const long long t0 = clock64();
                    mbar_wait2_poll(&bars->a_full, (na++) & 1, p.flag);
                    t_a += clock64() - t0;
                    tc_fence_after();
                }
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t a_hi = smem_base + E2_AHI + ks * E2_SLAB, a_lo = smem_base + E2_ALO + ks * E2_SLAB;
                    {
                        const uint32_t b = wait_unit();
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma2_f16(d, umma_desc(a_hi + k * 32, 16, ATOM_BYTES), umma_desc(b + k * 32, 16, ATOM_BYTES),
                                      IDESC2_N256, (accumulate || ks > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma2_f16(d, umma_desc(a_lo + k * 32, 16, ATOM_BYTES), umma_desc(b + k * 32, 16, ATOM_BYTES),
                                      IDESC2_N256, 1u);
                        umma2_commit(&bars->empty[g % E2_RING]);
                        ++g;
                    }
                    {
                        const uint32_t b = wait_unit();
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma2_f16(d, umma_desc(a_hi + k * 32, 16, ATOM_BYTES), umma_desc(b + k * 32, 16, ATOM_BYTES),
                                      IDESC2_N256, 1u);
                        umma2_commit(&bars->empty[g % E2_RING]);
                        ++g;
                    }
                }
            };
            if (p.do_q) {
                gemm(S0, false, true);  umma2_commit(&bars->s_full[0]);     // q
                gemm(S1, false, true);  umma2_commit(&bars->s_full[1]);     // msg
                gemm(S0, false, true);  umma2_commit(&bars->s_full[0]);     // h_a
                gemm(S1, false, false); umma2_commit(&bars->s_full[1]);     // h_b
                gemm(S0, false, true);  umma2_commit(&bars->s_full[0]);     // y  = gelu(h_a) W2a^T
                gemm(S0, true, true);   umma2_commit(&bars->s_full[0]);     // y += gelu(h_b) W2b^T
            }
            if (p.do_kv) {
                gemm(S0, false, true);      umma2_commit(&bars->s_full[0]); // v
                gemm(S1, false, dec_mode);  umma2_commit(&bars->s_full[1]); // k
                // KV rounds: round 0 = 64-channel groups 0,1 (N half 0), round 1 = groups 2,3.  V slabs live in the
                // operand image (group g: slab g of hi / lo), the round's Kf slabs in the (drained) weight ring.
                for (int round = 0; round < 2; ++round) {
                    mbar_wait2_poll(round == 0 ? &bars->kv_a : &bars->kv_b, 0, p.flag);
                    tc_fence_after();
                    for (int j = 0; j < 2; ++j) {
                        const int grp = round * 2 + j;
                        const uint32_t kf_hi = smem_base + E2_RINGOFF + j * E2_SLAB, kf_lo = kf_hi + 2 * E2_SLAB;
                        const uint32_t v_hi = smem_base + E2_AHI + grp * E2_SLAB, v_lo = smem_base + E2_ALO + grp * E2_SLAB;
                        const uint32_t dkv = (round == 0 ? S0 : S1) + j * 64;
#pragma unroll
                        for (int k = 0; k < HT / 16; ++k)
                            umma2_f16(dkv, umma_desc(kf_hi + k * 2048, E2_SLAB, ATOM_BYTES),
                                      umma_desc(v_hi + k * 2048, E2_SLAB, ATOM_BYTES), IDESC2_KV, k);
#pragma unroll
                        for (int k = 0; k < HT / 16; ++k)
                            umma2_f16(dkv, umma_desc(kf_lo + k * 2048, E2_SLAB, ATOM_BYTES),
                                      umma_desc(v_hi + k * 2048, E2_SLAB, ATOM_BYTES), IDESC2_KV, 1u);
#pragma unroll
                        for (int k = 0; k < HT / 16; ++k)
                            umma2_f16(dkv, umma_desc(kf_hi + k * 2048, E2_SLAB, ATOM_BYTES),
                                      umma_desc(v_lo + k * 2048, E2_SLAB, ATOM_BYTES), IDESC2_KV, 1u);
                    }
                    umma2_commit(&bars->s_full[round]);
                }
            }
            if (p.dbg_clock) {
                long long* o = p.dbg_clock + (size_t)tile * 4;
                o[0] = clock64() - t_begin; o[1] = t_a; o[2] = t_ring; o[3] = 0;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ row warps
        const int q = warp & 3, cq = warp >> 2;            // TMEM lane quarter; 64-column chunk inside the N half
        const int rh = q & 1, nh = q >> 1;                 // row half (rows 32*rh ..), N half (columns 128*nh ..)
        const int r = rh * 32 + lane;                      // token row inside this CTA
        const int rt = (int)rank * HT + r;                 // token row inside the 128-row tile
        const bool valid = rt < ti.valid;
        const int cbase = nh * 128 + cq * 64;              // this thread's 64 columns: [cbase, cbase + 64)
        const int idx = nh * 2 + cq;                       // 0..3: which quarter of the row
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t tcol = cq * 64;                     // TMEM column of cbase inside an accumulator
        float* X = reinterpret_cast<float*>(smem + E2_X);
        uint8_t* img_hi = smem + E2_AHI;
        uint8_t* img_lo = smem + E2_ALO;
        const float* post = (ti.set == 0 ? p.post1 : p.post2);
        const uint32_t a_full_addr = mapa_u32(smem_u32(&bars->a_full), 0);
        uint32_t ns0 = 0, ns1 = 0;
        auto wait_s = [&](int b) {
            mbar_wait2(&bars->s_full[b], (b ? ns1++ : ns0++) & 1, p.flag);
            tc_fence_after();
        };
        auto publish_to = [&](uint32_t cluster_addr) {     // this warp's part of the operand image is written
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(cluster_addr);
        };
        auto publish = [&]() { publish_to(a_full_addr); };
        // sum of one value per thread over the 4 threads that share a token row (3 slots of 64 floats)
        auto row_sum4 = [&](float part) -> float {
            if (idx != 0) X[(idx - 1) * HT + r] = part;
            named_bar_sync(1, E2_ROW_THREADS);
            float tot = part;
            if (idx == 0) { tot = (part + X[r]) + (X[HT + r] + X[2 * HT + r]); X[r] = tot; }
            named_bar_sync(1, E2_ROW_THREADS);
            if (idx != 0) tot = X[r];
            named_bar_sync(1, E2_ROW_THREADS);
            return tot;
        };
        // ---- residual stream of this thread: columns [cbase, cbase+64) of row rt, as two 32-column passes
        float x[2][32];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int c0 = cbase + pass * 32;
            if (p.load_feat) {
                const float* feat = ti.set == 0 ? p.feat1 : p.feat2;
                const float* f = feat + ((size_t)ti.b * C + c0) * ti.L + (size_t)ti.ti * TILE + rt;
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] = valid ? f[(size_t)e * ti.L] : 0.f;
            } else {
#pragma unroll
                for (int jq = 0; jq < 8; ++jq) {
                    const float4 v = *reinterpret_cast<const float4*>(p.xt + xt_off(tile, (c0 >> 2) + jq, rt));
                    x[pass][jq * 4 + 0] = v.x; x[pass][jq * 4 + 1] = v.y; x[pass][jq * 4 + 2] = v.z; x[pass][jq * 4 + 3] = v.w;
                }
            }
        }
        // operand image <- [LN](x) [+ pos] (two-pass statistics; gamma == nullptr: no LayerNorm), then publish
        auto image_from_x = [&](const float* __restrict__ gamma, const float* __restrict__ beta, bool with_pos) {
            float mean = 0.f, rstd = 1.f;
            if (gamma) {
                float s = 0.f;
#pragma unroll
                for (int e = 0; e < 32; ++e) s += x[0][e] + x[1][e];
                mean = row_sum4(s) * (1.f / C);
                float sq = 0.f;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float d0 = x[0][e] - mean, d1 = x[1][e] - mean;
                    sq = fmaf(d0, d0, sq);
                    sq = fmaf(d1, d1, sq);
                }
                rstd = rsqrtf(row_sum4(sq) * (1.f / C) + LN_EPS);
            }
            const float shift = -mean * rstd;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = cbase + pass * 32;
                float v[32];
#pragma unroll
                for (int jq = 0; jq < 8; ++jq) {
                    float4 ps = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (with_pos) ps = *reinterpret_cast<const float4*>(post + xt_off(ti.ti, (c0 >> 2) + jq, rt));
                    float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gamma) {
                        g4 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + jq);
                        b4 = __ldg(reinterpret_cast<const float4*>(beta + c0) + jq);
                    }
                    v[jq * 4 + 0] = fmaf(fmaf(x[pass][jq * 4 + 0], rstd, shift), g4.x, b4.x + ps.x);
                    v[jq * 4 + 1] = fmaf(fmaf(x[pass][jq * 4 + 1], rstd, shift), g4.y, b4.y + ps.y);
                    v[jq * 4 + 2] = fmaf(fmaf(x[pass][jq * 4 + 2], rstd, shift), g4.z, b4.z + ps.z);
                    v[jq * 4 + 3] = fmaf(fmaf(x[pass][jq * 4 + 3], rstd, shift), g4.w, b4.w + ps.w);
                }
                store_row32_split<E2_SLAB>(img_hi, img_lo, r, c0, v);
            }
            publish();
        };

        if (p.do_q) {
            // (E0) A = LNq(x) + pos
            image_from_x(p.lnq_g, p.lnq_b, true);
            // (E1) A = phi(q) / Z   (one head per 32-column pass)
            const float* ks = p.ksum + (size_t)src_img * C;
            wait_s(0);
            const float eps_s = ATTN_EPS / (float)src_len;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = cbase + pass * 32;
                float v[32];
                tmem_ld32(S0 + lane_addr + tcol + pass * 32, v);
                float den = 0.f;
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const float4 k4 = __ldg(reinterpret_cast<const float4*>(ks + c0) + e4);
                    v[e4 * 4 + 0] = elu1(v[e4 * 4 + 0]); den = fmaf(v[e4 * 4 + 0], k4.x, den);
                    v[e4 * 4 + 1] = elu1(v[e4 * 4 + 1]); den = fmaf(v[e4 * 4 + 1], k4.y, den);
                    v[e4 * 4 + 2] = elu1(v[e4 * 4 + 2]); den = fmaf(v[e4 * 4 + 2], k4.z, den);
                    v[e4 * 4 + 3] = elu1(v[e4 * 4 + 3]); den = fmaf(v[e4 * 4 + 3], k4.w, den);
                }
                const float inv = 1.f / (den + eps_s);
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] *= inv;
                store_row32_split<E2_SLAB>(img_hi, img_lo, r, c0, v);
            }
            publish();
            // (E2) x += msg ; A = LN2(x)
            wait_s(1);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                float v[32];
                tmem_ld32(S1 + lane_addr + tcol + pass * 32, v);
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] += v[e];
            }
            image_from_x(p.ln2_g, p.ln2_b, false);
            // (E3) A = gelu(h_a): the image is free once h_b (the second GEMM reading LN2(x)) has completed
            // (E4) A = gelu(h_b): the image is free once y = gelu(h_a) W2a^T has completed
#pragma unroll 1
            for (int which = 0; which < 2; ++which) {
                if (which == 0) wait_s(0);
                const uint32_t S = which ? S1 : S0;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const int c0 = cbase + pass * 32;
                    float v[32];
                    tmem_ld32(S + lane_addr + tcol + pass * 32, v);
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = gelu_erf(v[e]);
                    if (pass == 0) wait_s(which == 0 ? 1 : 0);
                    store_row32_split<E2_SLAB>(img_hi, img_lo, r, c0, v);
                }
                publish();      // the consuming GEMM starts (and may overwrite S0) only after all 16 warps got here
            }
            // (E5) x += y
            wait_s(0);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                float v[32];
                tmem_ld32(S0 + lane_addr + tcol + pass * 32, v);
#pragma unroll
                for (int e = 0; e < 32; ++e) x[pass][e] += v[e];
            }
            tc_fence_before();
        }
        if (p.store_x) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = cbase + pass * 32;
#pragma unroll
                for (int jq = 0; jq < 8; ++jq)
                    *reinterpret_cast<float4*>(p.xt + xt_off(tile, (c0 >> 2) + jq, rt)) =
                        make_float4(x[pass][jq * 4], x[pass][jq * 4 + 1], x[pass][jq * 4 + 2], x[pass][jq * 4 + 3]);
            }
        }
        if (p.do_kv) {
            if (!dec_mode) {
                image_from_x(p.lnkv_g, p.lnkv_b, true);
                wait_s(0);
                wait_s(1);
            } else {
                image_from_x(nullptr, nullptr, false);
                wait_s(0);
                image_from_x(nullptr, nullptr, true);
                wait_s(1);
            }
            // Both projections are complete: the operand image and the weight ring are free.
            float* part = p.kv_part + ((size_t)tile * 2 + rank) * KVS;
            // V slab of this thread's 64-channel group idx (MN-major: the 64 columns of a slab row are the channels)
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = cbase + pass * 32;
                float v[32];
                tmem_ld32(S0 + lane_addr + tcol + pass * 32, v);
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
                }
                store_row32_split<E2_SLAB>(img_hi + idx * E2_SLAB, img_lo + idx * E2_SLAB, r, pass * 32, v);
            }
            const uint32_t kv_a_addr = mapa_u32(smem_u32(&bars->kv_a), 0), kv_b_addr = mapa_u32(smem_u32(&bars->kv_b), 0);
            uint8_t* kf_hi = smem + E2_RINGOFF + cq * E2_SLAB;
            uint8_t* kf_lo = kf_hi + 2 * E2_SLAB;
            if (nh == 1) publish_to(kv_a_addr);            // V written and S0 read; Kf of N half 1 follows in round 1
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = cbase + pass * 32;
                float v[32];
                tmem_ld32(S1 + lane_addr + tcol + pass * 32, v);
                if (p.bk) {
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bk + c0) + e4);
                        v[e4 * 4] += b4.x; v[e4 * 4 + 1] += b4.y; v[e4 * 4 + 2] += b4.z; v[e4 * 4 + 3] += b4.w;
                    }
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = valid ? elu1(v[e]) : 0.f;
                // round 1 re-uses the Kf slabs of round 0: wait until the round-0 MMAs have completed
                if (nh == 1 && pass == 0) wait_s(0);
                store_row32_split<E2_SLAB>(kf_hi, kf_lo, r, pass * 32, v);
                // Ksum[c0 + j]: butterfly transpose-reduce over the warp's 32 rows, then across the two row halves
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool up = (lane & off) != 0;
#pragma unroll
                    for (int i = 0; i < off; ++i) {
                        const float send = up ? v[i] : v[i + off];
                        const float keep = up ? v[i + off] : v[i];
                        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                if (rh == 1) X[idx * 32 + lane] = v[0];
                if (nh == 0) named_bar_sync(2, 128); else named_bar_sync(3, 128);
                if (rh == 0) part[NH * HD * HD + c0 + lane] = v[0] + X[idx * 32 + lane];
                if (nh == 0) named_bar_sync(2, 128); else named_bar_sync(3, 128);
            }
            publish_to(nh == 0 ? kv_a_addr : kv_b_addr);
            // results: CTA `rank` keeps n-half `rank` of the pair product = TMEM lanes 64*rank ..: the warps with
            // nh == rank; warp (cq, rh) reads head 2*(2*round + cq) + rh
#pragma unroll 1
            for (int round = 0; round < 2; ++round) {
                if (!(nh == 1 && round == 0)) wait_s(round);        // N-half-1 warps consumed s_full[0] above
                if (nh == (int)rank) {
                    const int h = 2 * (2 * round + cq) + rh;
                    float v[32];
                    tmem_ld32((round == 0 ? S0 : S1) + lane_addr + cq * 64 + rh * 32, v);
                    float* o = part + h * HD * HD + lane * HD;
#pragma unroll
                    for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(o + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                }
            }
            tc_fence_before();
        }
    }
    // teardown: nobody leaves while the peer may still signal its barriers or the pair's MMAs read its memory
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == E2_WARP_PRODUCER) tmem_dealloc2(tmem, 256);
}

}  // namespace oetr
