// tcgen05 / TMEM / bulk-TMA kernels of the OETR encoder (OETR_PREC_FP16 path).
//
// Work decomposition: one CTA = one tile of 128 tokens of one image; 10 warps:
//   warps 0-7  "row" warps: thread <-> (token row, half of the 256 channels).  They own every row-wise step
//              (LayerNorm, +pos, elu+1, 1/Z, GELU, fp32->fp16 operand images) and read/write TMEM
//   warp 8     weight producer: streams pre-swizzled 16 KB fp16 weight chunks global->shared with cp.async.bulk
//              through a 4-stage mbarrier ring
//   warp 9     MMA issuer: one lane issues tcgen05.mma (M=128, fp16 x fp16 -> fp32 in TMEM)
// The fp32 residual stream of the tile lives in TMEM columns [0,256): the merge and MLP-down GEMMs accumulate
// straight into it (x += ...), so the residual adds cost nothing.  Columns [256,512) are the scratch accumulator.
//
// Two kernels per encoder layer (reference src/models/transformer.py:104-142, linear_attention.py:22-50):
//   k_tc_kv     source side: LN_kv(+pos) -> v,k projections -> elu+1 -> per-tile KV = K^T V and Ksum
//   k_tc_layer  query side : LN_q(+pos) -> q projection -> elu+1, Z -> Q.KV -> merge (+=x) -> LN2 -> MLP (+=x)
// Per-tile KV partials are summed by the consumer, which keeps the reduction order fixed (deterministic).
#include "tc_common.cuh"
#include "tc_path.cuh"

#include <cmath>
#include <cstdio>
#include <vector>

namespace oetr {
using namespace tc;

constexpr int TILE = 128;                    // tokens per CTA
constexpr int N_ROW_THREADS = 256;           // warps 0-7
constexpr int N_THREADS = 320;               // + producer warp + MMA warp
constexpr int CHUNK_HALFS = 128 * 64;        // one weight chunk: [128 N rows][64 K cols] fp16 = 16 KB
constexpr uint32_t CHUNK_BYTES = CHUNK_HALFS * 2;
constexpr int RING = 4;
constexpr uint32_t SLAB_BYTES = TILE * 128;  // one [128 x 64] fp16 operand slab = 16 KB
constexpr uint32_t AIMG_BYTES = 4 * SLAB_BYTES;   // a [128 x 256] fp16 operand image = 64 KB
constexpr int CHUNKS_PER_GEMM = 8;           // a 256x256 weight block = 2 N-halves x 4 K-slabs
constexpr int KV_STREAM_CHUNKS = 16;         // Wv | Wk
constexpr int LAYER_STREAM_CHUNKS = 48;      // Wq | Wm | W1a | W2a | W1b | W2b
constexpr size_t ENC_LAYER_HALFS = (size_t)(KV_STREAM_CHUNKS + LAYER_STREAM_CHUNKS) * CHUNK_HALFS;
constexpr size_t DEC_LAYER_HALFS = (size_t)KV_STREAM_CHUNKS * CHUNK_HALFS;

constexpr uint32_t IDESC_N128 = umma_idesc_f16(128, 128, 0, 0);
constexpr uint32_t IDESC_N32 = umma_idesc_f16(128, 32, 0, 0);
constexpr uint32_t IDESC_KV = umma_idesc_f16(128, 128, 1, 1);      // both operands MN-major (token = K)
constexpr uint32_t IDESC_KSUM = umma_idesc_f16(128, 16, 1, 1);

// shared-memory map (dynamic, 1024-byte aligned)
constexpr uint32_t SM_A0 = 0;
constexpr uint32_t SM_A1 = SM_A0 + AIMG_BYTES;
constexpr uint32_t SM_RING = SM_A1 + AIMG_BYTES;
constexpr uint32_t SM_AUX = SM_RING + RING * CHUNK_BYTES;          // 16 KB: ones slab (k_tc_kv) / KV^T image (k_tc_layer)
constexpr uint32_t SM_VEC = SM_AUX + SLAB_BYTES;                   // 8 float[256] vectors
constexpr uint32_t SM_RED = SM_VEC + 8 * 256 * 4;                  // float[2][2][128] LayerNorm partials
constexpr uint32_t SM_BAR = SM_RED + 2 * 2 * 128 * 4;              // mbarriers + tmem pointer
constexpr uint32_t SM_TOTAL = SM_BAR + 128;
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");

struct Bars {
    uint64_t full[RING], empty[RING];
    uint64_t a_full;      // row warps -> MMA: operand image(s) written (count 256)
    uint64_t s_full;      // MMA -> row warps: accumulator ready (tcgen05.commit)
    uint64_t s_full2;     // second accumulator of k_tc_kv (two commits may be in flight back to back)
    uint32_t tmem_base;
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------------------------
// weight images
// ---------------------------------------------------------------------------------------------------------
// 8 chunks (n-half outer, k-slab inner) of the 256x256 block W[row0.., col0..] (row-major fp32, leading dim ld)
__global__ void k_make_chunks(const float* __restrict__ W, int ld, int row0, int col0, __half* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // one 16-byte chunk (8 halfs) per thread
    if (idx >= CHUNKS_PER_GEMM * 128 * 8) return;
    const int j = idx & 7, r = (idx >> 3) & 127, chunk = idx >> 10;
    const int nh = chunk >> 2, ks = chunk & 3;
    const float* src = W + (size_t)(row0 + nh * 128 + r) * ld + col0 + ks * 64 + j * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = src[e];
    uint8_t* dst = reinterpret_cast<uint8_t*>(out + (size_t)chunk * CHUNK_HALFS) + slab_chunk_off(r, j);
    *reinterpret_cast<uint4*>(dst) = pack8_f16(v);
}

int tc_prepare_weights(const float* d_w, const WLayout& L, TcWeights& out, char* msg, size_t msg_len) {
    out.enc_layer_halfs = ENC_LAYER_HALFS;
    out.dec_layer_halfs = DEC_LAYER_HALFS;
    if (cudaMalloc(&out.enc_img, N_ENC * ENC_LAYER_HALFS * sizeof(__half)) != cudaSuccess ||
        cudaMalloc(&out.dec_img, N_DEC * DEC_LAYER_HALFS * sizeof(__half)) != cudaSuccess) {
        snprintf(msg, msg_len, "weight image allocation failed");
        return -1;
    }
    const int grid = (CHUNKS_PER_GEMM * 128 * 8 + 255) / 256;
    const size_t G = (size_t)CHUNKS_PER_GEMM * CHUNK_HALFS;
    for (int i = 0; i < N_ENC; ++i) {
        const EncW& e = L.enc[i];
        __half* o = out.enc_img + (size_t)i * ENC_LAYER_HALFS;
        k_make_chunks<<<grid, 256>>>(d_w + e.wv, C, 0, 0, o + 0 * G);           // kv stream: Wv, Wk
        k_make_chunks<<<grid, 256>>>(d_w + e.wk, C, 0, 0, o + 1 * G);
        k_make_chunks<<<grid, 256>>>(d_w + e.wq, C, 0, 0, o + 2 * G);           // layer stream
        k_make_chunks<<<grid, 256>>>(d_w + e.wm, C, 0, 0, o + 3 * G);
        k_make_chunks<<<grid, 256>>>(d_w + e.w1, C, 0, 0, o + 4 * G);           // W1[0:256, :]
        k_make_chunks<<<grid, 256>>>(d_w + e.w2, FF, 0, 0, o + 5 * G);          // W2[:, 0:256]
        k_make_chunks<<<grid, 256>>>(d_w + e.w1, C, 256, 0, o + 6 * G);         // W1[256:512, :]
        k_make_chunks<<<grid, 256>>>(d_w + e.w2, FF, 0, 256, o + 7 * G);        // W2[:, 256:512]
    }
    for (int j = 0; j < N_DEC; ++j) {
        const DecW& d = L.dec[j];
        __half* o = out.dec_img + (size_t)j * DEC_LAYER_HALFS;
        k_make_chunks<<<grid, 256>>>(d_w + d.ca.wv, C, 0, 0, o + 0 * G);
        k_make_chunks<<<grid, 256>>>(d_w + d.ca.wk, C, 0, 0, o + 1 * G);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(msg, msg_len, "weight image kernels: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

void tc_free_weights(TcWeights& w) {
    cudaFree(w.enc_img);
    cudaFree(w.dec_img);
    w.enc_img = w.dec_img = nullptr;
}

// ---------------------------------------------------------------------------------------------------------
// device building blocks shared by the kernels and the self-test
// ---------------------------------------------------------------------------------------------------------
struct RingState { uint32_t g = 0; };       // running chunk counter (producer and consumer each keep one)

__device__ __forceinline__ void produce_chunks(const __half* wimg, int nchunks, uint8_t* smem, Bars* bars, int* flag) {
    for (int g = 0; g < nchunks; ++g) {
        const int st = g % RING;
        mbar_wait(&bars->empty[st], ((g / RING) & 1) ^ 1, flag);
        mbar_arrive_expect_tx(&bars->full[st], CHUNK_BYTES);
        bulk_g2s(smem + SM_RING + st * CHUNK_BYTES, wimg + (size_t)g * CHUNK_HALFS, CHUNK_BYTES, &bars->full[st]);
    }
}

// D[128 x 256] (tmem columns d_col..d_col+255) (+)= A[128 x 256](image at a_off) . W^T, consuming 8 ring chunks
__device__ __forceinline__ void mma_gemm256(uint32_t smem_base, uint32_t a_off, uint32_t d_tmem, bool accumulate,
                                            RingState& rs, Bars* bars, int* flag) {
    for (int nh = 0; nh < 2; ++nh)
        for (int ks = 0; ks < 4; ++ks) {
            const int st = rs.g % RING;
            mbar_wait(&bars->full[st], (rs.g / RING) & 1, flag);
            tc_fence_after();
            const uint32_t a_slab = smem_base + a_off + ks * SLAB_BYTES;
            const uint32_t b_slab = smem_base + SM_RING + st * CHUNK_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_f16(d_tmem + nh * 128, umma_desc(a_slab + k * 32, 16, ATOM_BYTES),
                         umma_desc(b_slab + k * 32, 16, ATOM_BYTES), IDESC_N128, (accumulate || ks > 0 || k > 0) ? 1u : 0u);
            umma_commit(&bars->empty[st]);
            rs.g++;
        }
}

// per-head O_h[128 x 32] = Qf[:, 32h:32h+32] . KVt_h^T ; Qf image at a_off, KV^T image (4 slabs of 32 rows) at SM_AUX
__device__ __forceinline__ void mma_attn_apply(uint32_t smem_base, uint32_t a_off, uint32_t d_tmem) {
#pragma unroll
    for (int h = 0; h < NH; ++h) {
        const uint32_t a_addr = smem_base + a_off + (h >> 1) * SLAB_BYTES + (h & 1) * 64;
        const uint32_t b_addr = smem_base + SM_AUX + (h >> 1) * (32 * 128) + (h & 1) * 64;
#pragma unroll
        for (int k = 0; k < 2; ++k)
            umma_f16(d_tmem + h * 32, umma_desc(a_addr + k * 32, 16, ATOM_BYTES), umma_desc(b_addr + k * 32, 16, ATOM_BYTES),
                     IDESC_N32, k);
    }
}

// KV halves: D[0:128) = Kf[:,0:128]^T V[:,0:128], D[128:256) = Kf[:,128:256]^T V[:,128:256];
// Ksum: D[256:272) / D[272:288) = Kf-half^T . ones.   Kf image at SM_A0, V image at SM_A1, ones slab at SM_AUX.
// lbo/sbo are parameters only so that the self-test can probe the MN-major descriptor convention.
__device__ __forceinline__ void mma_kv(uint32_t smem_base, uint32_t d_tmem, uint32_t lbo, uint32_t sbo) {
    for (int half = 0; half < 2; ++half) {
        const uint32_t a0 = smem_base + SM_A0 + half * 2 * SLAB_BYTES;
        const uint32_t b0 = smem_base + SM_A1 + half * 2 * SLAB_BYTES;
#pragma unroll
        for (int k = 0; k < TILE / 16; ++k) {
            const uint64_t ad = umma_desc(a0 + k * 16 * 128, lbo, sbo);
            umma_f16(d_tmem + half * 128, ad, umma_desc(b0 + k * 16 * 128, lbo, sbo), IDESC_KV, k);
            umma_f16(d_tmem + 256 + half * 16, ad, umma_desc(smem_base + SM_AUX + k * 16 * 128, lbo, sbo), IDESC_KSUM, k);
        }
    }
}

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x + 1.f : __expf(x); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

// common prologue: barriers, TMEM allocation.  Returns the TMEM base address.
__device__ __forceinline__ uint32_t cta_setup(uint8_t* smem, Bars* bars) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        mbar_init(&bars->a_full, N_ROW_THREADS);
        mbar_init(&bars->s_full, 1);
        mbar_init(&bars->s_full2, 1);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return bars->tmem_base;
}
__device__ __forceinline__ void cta_teardown(uint32_t tmem_base) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 8) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------
// tile bookkeeping
// ---------------------------------------------------------------------------------------------------------
struct TileGeom {
    int B, L1, L2, T1, T2;          // tiles per image: T = ceil(L/128)
    __host__ __device__ int tiles() const { return B * (T1 + T2); }
};
struct TileInfo { int set, b, ti, L, T, img, valid, first_tile_of_img; };
__device__ __forceinline__ TileInfo tile_info(const TileGeom& g, int t) {
    TileInfo ti;
    if (t < g.B * g.T1) { ti.set = 0; ti.b = t / g.T1; ti.ti = t % g.T1; ti.L = g.L1; ti.T = g.T1; ti.first_tile_of_img = ti.b * g.T1; }
    else { const int u = t - g.B * g.T1; ti.set = 1; ti.b = u / g.T2; ti.ti = u % g.T2; ti.L = g.L2; ti.T = g.T2;
           ti.first_tile_of_img = g.B * g.T1 + ti.b * g.T2; }
    ti.img = ti.set * g.B + ti.b;
    ti.valid = min(TILE, ti.L - ti.ti * TILE);
    return ti;
}
// tile-blocked fp32 layout [tile][64 col-quads][128 rows][4]: a warp's rows read/write one col-quad coalesced
__device__ __forceinline__ size_t xt_off(int tile, int quad, int r) { return (((size_t)tile * 64 + quad) * TILE + r) * 4; }

// LayerNorm statistics of a row split over two threads (column halves): shifted one-pass sums, combined via smem
__device__ __forceinline__ void ln_combine(float s, float q, float* red, int r, int ch, float& mean_shifted, float& rstd) {
    red[(0 * 2 + ch) * TILE + r] = s;
    red[(1 * 2 + ch) * TILE + r] = q;
    named_bar_sync(1, N_ROW_THREADS);
    const float S = red[(0 * 2 + 0) * TILE + r] + red[(0 * 2 + 1) * TILE + r];
    const float Q = red[(1 * 2 + 0) * TILE + r] + red[(1 * 2 + 1) * TILE + r];
    mean_shifted = S * (1.f / C);
    rstd = rsqrtf(fmaxf(Q * (1.f / C) - mean_shifted * mean_shifted, 0.f) + LN_EPS);
    named_bar_sync(1, N_ROW_THREADS);          // red[] may be reused afterwards
}

// ---------------------------------------------------------------------------------------------------------
// k_tc_kv : source-side kernel
// ---------------------------------------------------------------------------------------------------------
struct KvParams {
    TileGeom g;
    const float* feat1;      // NCHW inputs (first encoder layer only) or nullptr
    const float* feat2;
    float* xt;               // tile-blocked residual stream (read when feat == nullptr, written when feat != nullptr)
    const float* post1;      // tile-blocked positional rows of set 0 / set 1
    const float* post2;
    const float *ln_g, *ln_b;   // LayerNorm affine or nullptr (decoder: memory is used un-normalised)
    const float *bk, *bv;       // projection biases or nullptr
    int pos_on_v;               // encoder: k and v share LN(x)+pos; decoder: k = x+pos, v = x
    const __half* wimg;         // 16 chunks: Wv | Wk
    float* kv_part;             // [tiles][KVS]
    int* flag;
    uint32_t mn_lbo, mn_sbo;
};

__global__ void __launch_bounds__(N_THREADS, 1) k_tc_kv(const KvParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    float* vec = reinterpret_cast<float*>(smem + SM_VEC);        // [0]=gamma [1]=beta [2]=bk [3]=bv
    float* red = reinterpret_cast<float*>(smem + SM_RED);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const TileInfo ti = tile_info(p.g, blockIdx.x);
    const uint32_t tmem = cta_setup(smem, bars);
    const uint32_t smem_base = smem_u32(smem);

    if (warp == 8) {
        if (lane == 0) produce_chunks(p.wimg, KV_STREAM_CHUNKS, smem, bars, p.flag);
        __syncwarp();
    } else if (warp == 9) {
        if (lane == 0) {
            RingState rs;
            mbar_wait(&bars->a_full, 0, p.flag);
            tc_fence_after();
            mma_gemm256(smem_base, p.pos_on_v ? SM_A0 : SM_A1, tmem + 0, false, rs, bars, p.flag);     // v
            umma_commit(&bars->s_full);
            mma_gemm256(smem_base, SM_A0, tmem + 256, false, rs, bars, p.flag);                         // k
            umma_commit(&bars->s_full2);
            mbar_wait(&bars->a_full, 1, p.flag);
            tc_fence_after();
            mma_kv(smem_base, tmem, p.mn_lbo, p.mn_sbo);
            umma_commit(&bars->s_full);
        }
        __syncwarp();
    } else {
        const int q = warp & 3, ch = warp >> 2;
        const int r = q * 32 + lane;
        const bool valid = r < ti.valid;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        // stage per-channel vectors and the ones slab
        vec[0 * 256 + tid] = p.ln_g ? p.ln_g[tid] : 1.f;
        vec[1 * 256 + tid] = p.ln_b ? p.ln_b[tid] : 0.f;
        vec[2 * 256 + tid] = p.bk ? p.bk[tid] : 0.f;
        vec[3 * 256 + tid] = p.bv ? p.bv[tid] : 0.f;
        {
            const __half2 one2 = __floats2half2_rn(1.f, 1.f);
            uint4 ones;
            ones.x = ones.y = ones.z = ones.w = *reinterpret_cast<const uint32_t*>(&one2);
            for (uint32_t i = tid; i < SLAB_BYTES / 16; i += N_ROW_THREADS) reinterpret_cast<uint4*>(smem + SM_AUX)[i] = ones;
        }
        const float* feat = ti.set == 0 ? p.feat1 : p.feat2;
        const float* post = (ti.set == 0 ? p.post1 : p.post2);
        const size_t tok0 = (size_t)ti.ti * TILE;
        auto load_quad = [&](int quad) -> float4 {                 // 4 channels [4*quad, 4*quad+4) of row r
            if (!valid) return make_float4(0.f, 0.f, 0.f, 0.f);
            if (feat) {
                const float* f = feat + ((size_t)ti.b * C + quad * 4) * ti.L + tok0 + r;
                return make_float4(f[0], f[(size_t)ti.L], f[(size_t)2 * ti.L], f[(size_t)3 * ti.L]);
            }
            return *reinterpret_cast<const float4*>(p.xt + xt_off(blockIdx.x, quad, r));
        };
        // pass 1: LayerNorm statistics (shifted by the row's first channel)
        float mean_s = 0.f, rstd = 1.f, shift = 0.f;
        if (p.ln_g) {
            shift = load_quad(0).x;
            float s = 0.f, sq = 0.f;
#pragma unroll 4
            for (int jq = 0; jq < 32; ++jq) {
                const float4 v = load_quad(ch * 32 + jq);
                const float a = v.x - shift, b = v.y - shift, c = v.z - shift, d = v.w - shift;
                s += (a + b) + (c + d);
                sq += (a * a + b * b) + (c * c + d * d);
            }
            ln_combine(s, sq, red, r, ch, mean_s, rstd);
        }
        __syncwarp();
        // pass 2: operand images.  A0 = LN(x)+pos (encoder) or x+pos (decoder); A1 = x (decoder only)
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = ch * 128 + cc * 32;
            float a0[32], a1[32];
#pragma unroll
            for (int jq = 0; jq < 8; ++jq) {
                const int quad = (c0 >> 2) + jq;
                const float4 v = load_quad(quad);
                if (feat) *reinterpret_cast<float4*>(p.xt + xt_off(blockIdx.x, quad, r)) = v;
                const float4 ps = *reinterpret_cast<const float4*>(post + xt_off(ti.ti, quad, r));
                const float x[4] = {v.x, v.y, v.z, v.w}, pp[4] = {ps.x, ps.y, ps.z, ps.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = c0 + jq * 4 + e;
                    const float n = p.ln_g ? ((x[e] - shift) - mean_s) * rstd * vec[c] + vec[256 + c] : x[e];
                    a0[jq * 4 + e] = n + pp[e];
                    a1[jq * 4 + e] = x[e];
                }
            }
            store_row32_f16(smem + SM_A0, SLAB_BYTES, r, c0, a0);
            if (!p.pos_on_v) store_row32_f16(smem + SM_A1, SLAB_BYTES, r, c0, a1);
        }
        fence_async_smem();
        mbar_arrive(&bars->a_full);
        // v = S0 + bv  -> V image (A1); padded rows are zero so they add nothing to KV
        mbar_wait(&bars->s_full, 0, p.flag);
        tc_fence_after();
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = ch * 128 + cc * 32;
            float v[32];
            tmem_ld32(lane_addr + c0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = valid ? v[e] + vec[3 * 256 + c0 + e] : 0.f;
            store_row32_f16(smem + SM_A1, SLAB_BYTES, r, c0, v);
        }
        // k = elu(S1 + bk) + 1 -> Kf image (A0)
        mbar_wait(&bars->s_full2, 0, p.flag);
        tc_fence_after();
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = ch * 128 + cc * 32;
            float v[32];
            tmem_ld32(lane_addr + 256 + c0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = valid ? elu1(v[e] + vec[2 * 256 + c0 + e]) : 0.f;
            store_row32_f16(smem + SM_A0, SLAB_BYTES, r, c0, v);
        }
        tc_fence_before();
        fence_async_smem();
        mbar_arrive(&bars->a_full);
        // KV diagonal blocks: this warp's TMEM lanes are the d-channels of head (4*ch + q)
        mbar_wait(&bars->s_full, 1, p.flag);
        tc_fence_after();
        {
            const int h = ch * 4 + q;
            float v[32];
            tmem_ld32(lane_addr + ch * 128 + q * 32, v);
            float* o = p.kv_part + (size_t)blockIdx.x * KVS + h * HD * HD + lane * HD;
#pragma unroll
            for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(o + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
            if (ch == 0) {       // Ksum: column 0 of the two ones-products
                tmem_ld32(lane_addr + 256, v);
                float* ks = p.kv_part + (size_t)blockIdx.x * KVS + NH * HD * HD;
                ks[q * 32 + lane] = v[0];
                ks[128 + q * 32 + lane] = v[16];
            }
        }
    }
    cta_teardown(tmem);
}

// ---------------------------------------------------------------------------------------------------------
// k_tc_layer : query-side kernel (one full encoder layer update of the tile's residual stream)
// ---------------------------------------------------------------------------------------------------------
struct LayerParams {
    TileGeom g;
    float* xt;
    const float *post1, *post2;
    const float *lnq_g, *lnq_b, *ln2_g, *ln2_b;
    const __half* wimg;         // 48 chunks: Wq | Wm | W1a | W2a | W1b | W2b
    const float* kv_part;       // per-tile partial summaries written by k_tc_kv
    int cross;                  // 1: read the partner image's summary (transformer.py:354-358)
    int* flag;
};

__global__ void __launch_bounds__(N_THREADS, 1) k_tc_layer(const LayerParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    float* vec = reinterpret_cast<float*>(smem + SM_VEC);   // [0]=lnq_g [1]=lnq_b [2]=ln2_g [3]=ln2_b [4]=Ksum
    float* red = reinterpret_cast<float*>(smem + SM_RED);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const TileInfo ti = tile_info(p.g, blockIdx.x);
    const uint32_t tmem = cta_setup(smem, bars);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t X = tmem, S = tmem + 256;

    if (warp == 8) {
        if (lane == 0) produce_chunks(p.wimg, LAYER_STREAM_CHUNKS, smem, bars, p.flag);
        __syncwarp();
    } else if (warp == 9) {
        if (lane == 0) {
            RingState rs;
            mbar_wait(&bars->a_full, 0, p.flag); tc_fence_after();
            mma_gemm256(smem_base, SM_A0, S, false, rs, bars, p.flag);         // q = (LNq(x)+pos) Wq^T
            umma_commit(&bars->s_full);
            mbar_wait(&bars->a_full, 1, p.flag); tc_fence_after();
            mma_attn_apply(smem_base, SM_A1, S);                               // O_h = Qf_h KV_h
            umma_commit(&bars->s_full);
            mbar_wait(&bars->a_full, 0, p.flag); tc_fence_after();
            mma_gemm256(smem_base, SM_A0, X, true, rs, bars, p.flag);          // x += (O/Z) Wm^T
            umma_commit(&bars->s_full);
            mbar_wait(&bars->a_full, 1, p.flag); tc_fence_after();
            mma_gemm256(smem_base, SM_A1, S, false, rs, bars, p.flag);         // h_a = LN2(x) W1a^T
            umma_commit(&bars->s_full);
            mbar_wait(&bars->a_full, 0, p.flag); tc_fence_after();
            mma_gemm256(smem_base, SM_A0, X, true, rs, bars, p.flag);          // x += gelu(h_a) W2a^T
            mma_gemm256(smem_base, SM_A1, S, false, rs, bars, p.flag);         // h_b = LN2(x) W1b^T
            umma_commit(&bars->s_full);
            mbar_wait(&bars->a_full, 1, p.flag); tc_fence_after();
            mma_gemm256(smem_base, SM_A0, X, true, rs, bars, p.flag);          // x += gelu(h_b) W2b^T
            umma_commit(&bars->s_full);
        }
        __syncwarp();
    } else {
        const int q = warp & 3, ch = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        vec[0 * 256 + tid] = p.lnq_g[tid]; vec[1 * 256 + tid] = p.lnq_b[tid];
        vec[2 * 256 + tid] = p.ln2_g[tid]; vec[3 * 256 + tid] = p.ln2_b[tid];
        // ---- linear-attention summary of the source image: sum the per-tile partials in tile order
        {
            int src_first = ti.first_tile_of_img, src_T = ti.T;
            if (p.cross) {
                if (ti.set == 0) { src_first = p.g.B * p.g.T1 + ti.b * p.g.T2; src_T = p.g.T2; }
                else { src_first = ti.b * p.g.T1; src_T = p.g.T1; }
            }
            const float* part = p.kv_part + (size_t)src_first * KVS;
            // KV[h][d][e] -> B operand rows e, K = d:  image slab h/2, row e, column (h&1)*32 + d
            for (int idx = tid; idx < NH * HD * HD; idx += N_ROW_THREADS) {
                float acc = 0.f;
                for (int t = 0; t < src_T; ++t) acc += part[(size_t)t * KVS + idx];
                const int h = idx >> 10, d = (idx >> 5) & 31, e = idx & 31;
                const uint32_t col = (h & 1) * 32 + d;
                uint8_t* dst = smem + SM_AUX + (h >> 1) * (32 * 128) + slab_chunk_off(e, col >> 3) + (col & 7) * 2;
                *reinterpret_cast<__half*>(dst) = __float2half_rn(acc);
            }
            float acc = 0.f;
            for (int t = 0; t < src_T; ++t) acc += part[(size_t)t * KVS + NH * HD * HD + tid];
            vec[4 * 256 + tid] = acc;
        }
        const float* post = (ti.set == 0 ? p.post1 : p.post2);
        // ---- (1) x -> TMEM, LNq statistics
        float mean_s, rstd, shift;
        {
            shift = p.xt[xt_off(blockIdx.x, 0, r)];
            float s = 0.f, sq = 0.f;
            for (int cc = 0; cc < 4; ++cc) {
                const int c0 = ch * 128 + cc * 32;
                float v[32];
#pragma unroll
                for (int jq = 0; jq < 8; ++jq) {
                    const float4 x4 = *reinterpret_cast<const float4*>(p.xt + xt_off(blockIdx.x, (c0 >> 2) + jq, r));
                    v[jq * 4 + 0] = x4.x; v[jq * 4 + 1] = x4.y; v[jq * 4 + 2] = x4.z; v[jq * 4 + 3] = x4.w;
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) { const float d = v[e] - shift; s += d; sq = fmaf(d, d, sq); }
                tmem_st32(X + lane_addr + c0, v);
            }
            tmem_st_wait();
            ln_combine(s, sq, red, r, ch, mean_s, rstd);
        }
        // ---- (2) A0 = LNq(x) + pos
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = ch * 128 + cc * 32;
            float v[32];
            tmem_ld32(X + lane_addr + c0, v);
#pragma unroll
            for (int jq = 0; jq < 8; ++jq) {
                const float4 ps = *reinterpret_cast<const float4*>(post + xt_off(ti.ti, (c0 >> 2) + jq, r));
                const float pp[4] = {ps.x, ps.y, ps.z, ps.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = c0 + jq * 4 + e;
                    v[jq * 4 + e] = ((v[jq * 4 + e] - shift) - mean_s) * rstd * vec[c] + vec[256 + c] + pp[e];
                }
            }
            store_row32_f16(smem + SM_A0, SLAB_BYTES, r, c0, v);
        }
        tc_fence_before(); fence_async_smem(); mbar_arrive(&bars->a_full);
        // ---- (3) Qf = elu(q)+1 -> A1 ; 1/Z per head (linear_attention.py:46)
        float inv_den[4];
        mbar_wait(&bars->s_full, 0, p.flag); tc_fence_after();
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = ch * 128 + cc * 32;            // one head per 32-column chunk
            float v[32];
            tmem_ld32(S + lane_addr + c0, v);
            float den = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) { v[e] = elu1(v[e]); den = fmaf(v[e], vec[4 * 256 + c0 + e], den); }
            inv_den[cc] = 1.f / (den + ATTN_EPS);
            store_row32_f16(smem + SM_A1, SLAB_BYTES, r, c0, v);
        }
        tc_fence_before(); fence_async_smem(); mbar_arrive(&bars->a_full);
        // ---- (4) message = (Qf KV) / Z -> A0
        mbar_wait(&bars->s_full, 1, p.flag); tc_fence_after();
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = ch * 128 + cc * 32;
            float v[32];
            tmem_ld32(S + lane_addr + c0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] *= inv_den[cc];
            store_row32_f16(smem + SM_A0, SLAB_BYTES, r, c0, v);
        }
        tc_fence_before(); fence_async_smem(); mbar_arrive(&bars->a_full);
        // ---- (5) after x += merge(message): A1 = LN2(x)
        mbar_wait(&bars->s_full, 0, p.flag); tc_fence_after();
        {
            float v[32];
            tmem_ld32(X + lane_addr, v);              // column 0 of the row = shift (both halves use the same)
            shift = v[0];
            float s = 0.f, sq = 0.f;
            for (int cc = 0; cc < 4; ++cc) {
                tmem_ld32(X + lane_addr + ch * 128 + cc * 32, v);
#pragma unroll
                for (int e = 0; e < 32; ++e) { const float d = v[e] - shift; s += d; sq = fmaf(d, d, sq); }
            }
            ln_combine(s, sq, red, r, ch, mean_s, rstd);
            for (int cc = 0; cc < 4; ++cc) {
                const int c0 = ch * 128 + cc * 32;
                tmem_ld32(X + lane_addr + c0, v);
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    v[e] = ((v[e] - shift) - mean_s) * rstd * vec[2 * 256 + c0 + e] + vec[3 * 256 + c0 + e];
                store_row32_f16(smem + SM_A1, SLAB_BYTES, r, c0, v);
            }
        }
        tc_fence_before(); fence_async_smem(); mbar_arrive(&bars->a_full);
        // ---- (6),(7) gelu(h) -> A0, twice (hidden halves)
        for (int half = 0; half < 2; ++half) {
            mbar_wait(&bars->s_full, half ? 0 : 1, p.flag); tc_fence_after();
            for (int cc = 0; cc < 4; ++cc) {
                const int c0 = ch * 128 + cc * 32;
                float v[32];
                tmem_ld32(S + lane_addr + c0, v);
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = gelu_erf(v[e]);
                store_row32_f16(smem + SM_A0, SLAB_BYTES, r, c0, v);
            }
            tc_fence_before(); fence_async_smem(); mbar_arrive(&bars->a_full);
        }
        // ---- (8) write the updated residual stream back
        mbar_wait(&bars->s_full, 1, p.flag); tc_fence_after();
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = ch * 128 + cc * 32;
            float v[32];
            tmem_ld32(X + lane_addr + c0, v);
#pragma unroll
            for (int jq = 0; jq < 8; ++jq)
                *reinterpret_cast<float4*>(p.xt + xt_off(blockIdx.x, (c0 >> 2) + jq, r)) =
                    make_float4(v[jq * 4], v[jq * 4 + 1], v[jq * 4 + 2], v[jq * 4 + 3]);
        }
    }
    cta_teardown(tmem);
}

// ---------------------------------------------------------------------------------------------------------
// small layout kernels
// ---------------------------------------------------------------------------------------------------------
// tile-blocked positional rows of an (hf,wf) map from the channel-last PE table [max_h][max_w][256]
__global__ void k_pos_tiles(const float* __restrict__ pe, int max_w, int wf, int L, float* __restrict__ post) {
    const int tile = blockIdx.x;
    for (int i = threadIdx.x; i < 64 * TILE; i += blockDim.x) {
        const int quad = i / TILE, r = i % TILE;
        const int l = tile * TILE + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l < L) v = reinterpret_cast<const float4*>(pe)[((size_t)(l / wf) * max_w + (l % wf)) * 64 + quad];
        *reinterpret_cast<float4*>(post + xt_off(tile, quad, r)) = v;
    }
}
// tile-blocked -> token-major [rows][256] (encoder output "memory")
__global__ void k_untile(const float* __restrict__ xt, TileGeom g, float* __restrict__ X) {
    const TileInfo ti = tile_info(g, blockIdx.x);
    const size_t row0 = (ti.set == 0 ? (size_t)ti.b * g.L1 : (size_t)g.B * g.L1 + (size_t)ti.b * g.L2) + (size_t)ti.ti * TILE;
    for (int i = threadIdx.x; i < ti.valid * 64; i += blockDim.x) {
        const int r = i >> 6, quad = i & 63;
        reinterpret_cast<float4*>(X)[(row0 + r) * 64 + quad] =
            *reinterpret_cast<const float4*>(xt + xt_off(blockIdx.x, quad, r));
    }
}
// per-image sum of per-tile partial summaries
__global__ void k_sum_partials(const float* __restrict__ part, TileGeom g, float* __restrict__ out) {
    const int img = blockIdx.x;                      // 0..2B-1
    const int set = img / g.B, b = img % g.B;
    const int T = set == 0 ? g.T1 : g.T2;
    const int first = set == 0 ? b * g.T1 : g.B * g.T1 + b * g.T2;
    for (int i = threadIdx.x; i < KVS; i += blockDim.x) {
        float acc = 0.f;
        for (int t = 0; t < T; ++t) acc += part[(size_t)(first + t) * KVS + i];
        out[(size_t)img * KVS + i] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static TileGeom make_geom(int B, int L1, int L2) {
    TileGeom g;
    g.B = B; g.L1 = L1; g.L2 = L2; g.T1 = (L1 + TILE - 1) / TILE; g.T2 = (L2 + TILE - 1) / TILE;
    return g;
}

void tc_carve(size_t& off, void* base, int B, int L1, int L2, TcWorkspace& w) {
    const TileGeom g = make_geom(B, L1, L2);
    char* b = static_cast<char*>(base);
    auto take = [&](size_t nfloats) {
        off = (off + 255) & ~size_t(255);
        float* p = b ? reinterpret_cast<float*>(b + off) : nullptr;
        off += nfloats * sizeof(float);
        return p;
    };
    w.xt = take((size_t)g.tiles() * TILE * C);
    w.post = take((size_t)(g.T1 + g.T2) * TILE * C);
    w.kv_part = take((size_t)g.tiles() * KVS);
    w.dec_kvs = take((size_t)N_DEC * 2 * B * KVS);
}

static bool g_attr_set = false;
static int set_attrs(char* msg, size_t msg_len) {
    if (g_attr_set) return 0;
    cudaError_t e1 = cudaFuncSetAttribute(k_tc_kv, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    cudaError_t e2 = cudaFuncSetAttribute(k_tc_layer, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        snprintf(msg, msg_len, "cudaFuncSetAttribute(max dynamic smem %u): %s", SM_TOTAL,
                 cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        return -1;
    }
    g_attr_set = true;
    return 0;
}

int tc_encoder(const TcWeights& tw, const float* d_w, const WLayout& L, const TcWorkspace& ws, const float* feat1,
               const float* feat2, int B, int hf1, int wf1, int hf2, int wf2, const float* d_pe, int max_w,
               float* X_out, int* flag, KernelProfiler* prof, cudaStream_t s, LaunchCounter& lc, char* msg, size_t msg_len) {
    if (set_attrs(msg, msg_len)) return -1;
    const int L1 = hf1 * wf1, L2 = hf2 * wf2;
    const TileGeom g = make_geom(B, L1, L2);
    float* post1 = ws.post;
    float* post2 = ws.post + (size_t)g.T1 * TILE * C;
    k_pos_tiles<<<g.T1, 256, 0, s>>>(d_pe, max_w, wf1, L1, post1); lc.n++;
    k_pos_tiles<<<g.T2, 256, 0, s>>>(d_pe, max_w, wf2, L2, post2); lc.n++;
    const int tiles = g.tiles();
    for (int i = 0; i < N_ENC; ++i) {
        const EncW& e = L.enc[i];
        const __half* img = tw.enc_img + (size_t)i * ENC_LAYER_HALFS;
        KvParams kp{};
        kp.g = g; kp.feat1 = i == 0 ? feat1 : nullptr; kp.feat2 = i == 0 ? feat2 : nullptr; kp.xt = ws.xt;
        kp.post1 = post1; kp.post2 = post2; kp.ln_g = d_w + e.lnkv_g; kp.ln_b = d_w + e.lnkv_b;
        kp.bk = nullptr; kp.bv = nullptr; kp.pos_on_v = 1; kp.wimg = img; kp.kv_part = ws.kv_part; kp.flag = flag;
        kp.mn_lbo = SLAB_BYTES; kp.mn_sbo = ATOM_BYTES;
        k_tc_kv<<<tiles, N_THREADS, SM_TOTAL, s>>>(kp); lc.n++;
        LayerParams lp{};
        lp.g = g; lp.xt = ws.xt; lp.post1 = post1; lp.post2 = post2;
        lp.lnq_g = d_w + e.lnq_g; lp.lnq_b = d_w + e.lnq_b; lp.ln2_g = d_w + e.ln2_g; lp.ln2_b = d_w + e.ln2_b;
        lp.wimg = img + (size_t)KV_STREAM_CHUNKS * CHUNK_HALFS; lp.kv_part = ws.kv_part; lp.cross = i & 1; lp.flag = flag;
        if (prof) prof->mark(s);
        k_tc_layer<<<tiles, N_THREADS, SM_TOTAL, s>>>(lp); lc.n++;
        if (prof) prof->mark(s);
    }
    // decoder cross-attention summaries: k = (memory+pos) Wk^T + bk, v = memory Wv^T + bv (transformer.py:243-249)
    for (int j = 0; j < N_DEC; ++j) {
        const DecW& d = L.dec[j];
        KvParams kp{};
        kp.g = g; kp.feat1 = nullptr; kp.feat2 = nullptr; kp.xt = ws.xt; kp.post1 = post1; kp.post2 = post2;
        kp.ln_g = nullptr; kp.ln_b = nullptr; kp.bk = d_w + d.ca.bk; kp.bv = d_w + d.ca.bv; kp.pos_on_v = 0;
        kp.wimg = tw.dec_img + (size_t)j * DEC_LAYER_HALFS; kp.kv_part = ws.kv_part; kp.flag = flag;
        kp.mn_lbo = SLAB_BYTES; kp.mn_sbo = ATOM_BYTES;
        k_tc_kv<<<tiles, N_THREADS, SM_TOTAL, s>>>(kp); lc.n++;
        k_sum_partials<<<2 * B, 256, 0, s>>>(ws.kv_part, g, ws.dec_kvs + (size_t)j * 2 * B * KVS); lc.n++;
    }
    k_untile<<<tiles, 256, 0, s>>>(ws.xt, g, X_out); lc.n++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(msg, msg_len, "tcgen05 encoder launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// self-test of the building blocks on one tile
// ---------------------------------------------------------------------------------------------------------
// Runs k_tc_kv and k_tc_layer on ONE synthetic 128-token image with LayerNorm disabled / identity so that the
// expected outputs are plain matrix products computed on the host:
//   errs[0] v-projection path + KV summary (K-major GEMM, MN-major KV MMA)   [with LBO = slab stride]
//   errs[1] Ksum (N=16 ones-product)
//   errs[2] unused (0)
//   errs[3] full layer kernel vs host fp32 emulation of the same tile
//   errs[4] timeout flag raised by any mbarrier wait (0 = none)
int tc_selftest(float* errs, int n_errs, char* msg, size_t msg_len) {
    for (int i = 0; i < n_errs; ++i) errs[i] = -1.f;
    if (set_attrs(msg, msg_len)) return -1;
    const int L = 100;                              // one partial tile: rows 100..127 are padding
    const TileGeom g = make_geom(1, L, L);
    auto frand = [](uint32_t& s) { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.f - 0.5f; };
    uint32_t seed = 12345u;
    std::vector<float> feat((size_t)2 * C * L), wk(C * C), wv(C * C), wq(C * C), wm(C * C), w1(FF * C), w2(C * FF);
    for (auto& v : feat) v = 2.f * frand(seed);
    for (auto* w : {&wk, &wv, &wq, &wm, &w1, &w2}) for (auto& v : *w) v = 0.25f * frand(seed);
    std::vector<float> ones(C, 1.f), zeros(C, 0.f), pos((size_t)2 * TILE * C, 0.f);
    auto h16 = [](float x) { return __half2float(__float2half_rn(x)); };

    float *d_feat, *d_wk, *d_wv, *d_wq, *d_wm, *d_w1, *d_w2, *d_ones, *d_zeros, *d_xt, *d_post, *d_part;
    __half* d_img;
    int* d_flag;
    const size_t img_halfs = ENC_LAYER_HALFS;
#define ST(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(msg, msg_len, "%s: %s", #call, cudaGetErrorString(e_)); return -1; } } while (0)
    ST(cudaMalloc(&d_feat, feat.size() * 4)); ST(cudaMalloc(&d_wk, wk.size() * 4)); ST(cudaMalloc(&d_wv, wv.size() * 4));
    ST(cudaMalloc(&d_wq, wq.size() * 4)); ST(cudaMalloc(&d_wm, wm.size() * 4)); ST(cudaMalloc(&d_w1, w1.size() * 4));
    ST(cudaMalloc(&d_w2, w2.size() * 4)); ST(cudaMalloc(&d_ones, C * 4)); ST(cudaMalloc(&d_zeros, C * 4));
    ST(cudaMalloc(&d_xt, (size_t)2 * TILE * C * 4)); ST(cudaMalloc(&d_post, pos.size() * 4));
    ST(cudaMalloc(&d_part, (size_t)2 * KVS * 4)); ST(cudaMalloc(&d_img, img_halfs * 2)); ST(cudaMalloc(&d_flag, 4));
    ST(cudaMemcpy(d_feat, feat.data(), feat.size() * 4, cudaMemcpyHostToDevice));
    ST(cudaMemcpy(d_wk, wk.data(), wk.size() * 4, cudaMemcpyHostToDevice));
    ST(cudaMemcpy(d_wv, wv.data(), wv.size() * 4, cudaMemcpyHostToDevice));
    ST(cudaMemcpy(d_wq, wq.data(), wq.size() * 4, cudaMemcpyHostToDevice));
    ST(cudaMemcpy(d_wm, wm.data(), wm.size() * 4, cudaMemcpyHostToDevice));
    ST(cudaMemcpy(d_w1, w1.data(), w1.size() * 4, cudaMemcpyHostToDevice));
    ST(cudaMemcpy(d_w2, w2.data(), w2.size() * 4, cudaMemcpyHostToDevice));
    ST(cudaMemcpy(d_ones, ones.data(), C * 4, cudaMemcpyHostToDevice));
    ST(cudaMemset(d_zeros, 0, C * 4)); ST(cudaMemset(d_post, 0, pos.size() * 4)); ST(cudaMemset(d_flag, 0, 4));
    const int grid = (CHUNKS_PER_GEMM * 128 * 8 + 255) / 256;
    const size_t G = (size_t)CHUNKS_PER_GEMM * CHUNK_HALFS;
    k_make_chunks<<<grid, 256>>>(d_wv, C, 0, 0, d_img + 0 * G);
    k_make_chunks<<<grid, 256>>>(d_wk, C, 0, 0, d_img + 1 * G);
    k_make_chunks<<<grid, 256>>>(d_wq, C, 0, 0, d_img + 2 * G);
    k_make_chunks<<<grid, 256>>>(d_wm, C, 0, 0, d_img + 3 * G);
    k_make_chunks<<<grid, 256>>>(d_w1, C, 0, 0, d_img + 4 * G);
    k_make_chunks<<<grid, 256>>>(d_w2, FF, 0, 0, d_img + 5 * G);
    k_make_chunks<<<grid, 256>>>(d_w1, C, 256, 0, d_img + 6 * G);
    k_make_chunks<<<grid, 256>>>(d_w2, FF, 0, 256, d_img + 7 * G);

    // host reference for the kv kernel with LN disabled and no pos: a = fp16(x); v = a Wv^T ; k = elu1(a Wk^T)
    std::vector<float> xa((size_t)2 * L * C);                      // [img][l][c]
    for (int im = 0; im < 2; ++im) for (int c = 0; c < C; ++c) for (int l = 0; l < L; ++l)
        xa[((size_t)im * L + l) * C + c] = feat[((size_t)im * C + c) * L + l];
    auto gemm_nt_h = [&](const std::vector<float>& A, int rows, int K, const std::vector<float>& W, int N, std::vector<double>& out) {
        out.assign((size_t)rows * N, 0.0);
        for (int i = 0; i < rows; ++i) for (int n = 0; n < N; ++n) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += (double)h16(A[(size_t)i * K + k]) * (double)h16(W[(size_t)n * K + k]);
            out[(size_t)i * N + n] = acc;
        }
    };
    std::vector<double> vh, kh;
    gemm_nt_h(xa, 2 * L, C, wv, C, vh);
    gemm_nt_h(xa, 2 * L, C, wk, C, kh);
    std::vector<double> kv_ref((size_t)2 * KVS, 0.0);
    for (int im = 0; im < 2; ++im) for (int l = 0; l < L; ++l) for (int h = 0; h < NH; ++h) for (int d = 0; d < HD; ++d) {
        const double kf = h16((float)(kh[((size_t)im * L + l) * C + h * HD + d] > 0 ? kh[((size_t)im * L + l) * C + h * HD + d] + 1.0
                                                                               : std::exp(kh[((size_t)im * L + l) * C + h * HD + d])));
        kv_ref[(size_t)im * KVS + NH * HD * HD + h * HD + d] += kf;
        for (int e = 0; e < HD; ++e)
            kv_ref[(size_t)im * KVS + h * HD * HD + d * HD + e] += kf * (double)h16((float)vh[((size_t)im * L + l) * C + h * HD + e]);
    }
    std::vector<float> part((size_t)2 * KVS);
    for (int variant = 0; variant < 1; ++variant) {   // (the swapped LBO/SBO probe addressed out of range on B200: convention settled)
        KvParams kp{};
        kp.g = g; kp.feat1 = d_feat; kp.feat2 = d_feat + (size_t)C * L; kp.xt = d_xt; kp.post1 = d_post; kp.post2 = d_post;
        kp.ln_g = nullptr; kp.ln_b = nullptr; kp.bk = nullptr; kp.bv = nullptr; kp.pos_on_v = 1; kp.wimg = d_img;
        kp.kv_part = d_part; kp.flag = d_flag;
        kp.mn_lbo = variant == 0 ? SLAB_BYTES : ATOM_BYTES; kp.mn_sbo = variant == 0 ? ATOM_BYTES : SLAB_BYTES;
        ST(cudaMemset(d_part, 0, part.size() * 4));
        k_tc_kv<<<2, N_THREADS, SM_TOTAL>>>(kp);
        ST(cudaDeviceSynchronize());
        ST(cudaMemcpy(part.data(), d_part, part.size() * 4, cudaMemcpyDeviceToHost));
        double e_kv = 0, e_ks = 0;
        for (int im = 0; im < 2; ++im) {
            for (int i = 0; i < NH * HD * HD; ++i) e_kv = std::max(e_kv, std::abs(part[(size_t)im * KVS + i] - kv_ref[(size_t)im * KVS + i]));
            for (int i = 0; i < NH * HD; ++i) e_ks = std::max(e_ks, std::abs(part[(size_t)im * KVS + NH * HD * HD + i] - kv_ref[(size_t)im * KVS + NH * HD * HD + i]));
        }
        double scale = 0;
        for (int i = 0; i < NH * HD * HD; ++i) scale = std::max(scale, std::abs(kv_ref[i]));
        if (variant == 0) { errs[0] = (float)(e_kv / scale); if (n_errs > 1) errs[1] = (float)(e_ks / L); }
    }
    if (n_errs > 2) errs[2] = 0.f;
    // layer kernel (self layer) with the variant-0 summaries: host emulation with the same rounding points
    if (n_errs > 3) {
        KvParams kp{};
        kp.g = g; kp.feat1 = d_feat; kp.feat2 = d_feat + (size_t)C * L; kp.xt = d_xt; kp.post1 = d_post; kp.post2 = d_post;
        kp.pos_on_v = 1; kp.wimg = d_img; kp.kv_part = d_part; kp.flag = d_flag; kp.mn_lbo = SLAB_BYTES; kp.mn_sbo = ATOM_BYTES;
        k_tc_kv<<<2, N_THREADS, SM_TOTAL>>>(kp);
        LayerParams lp{};
        lp.g = g; lp.xt = d_xt; lp.post1 = d_post; lp.post2 = d_post; lp.lnq_g = d_ones; lp.lnq_b = d_zeros;
        lp.ln2_g = d_ones; lp.ln2_b = d_zeros; lp.wimg = d_img + (size_t)KV_STREAM_CHUNKS * CHUNK_HALFS; lp.kv_part = d_part;
        lp.cross = 0; lp.flag = d_flag;
        k_tc_layer<<<2, N_THREADS, SM_TOTAL>>>(lp);
        ST(cudaDeviceSynchronize());
        std::vector<float> xt((size_t)2 * TILE * C);
        ST(cudaMemcpy(xt.data(), d_xt, xt.size() * 4, cudaMemcpyDeviceToHost));
        ST(cudaMemcpy(part.data(), d_part, part.size() * 4, cudaMemcpyDeviceToHost));
        double err = 0, scale = 0;
        const int im = 0;
        for (int l = 0; l < L; l += 7) {                 // a sample of rows of image 0
            std::vector<double> x(C), a(C), qv(C), msg_(C), hid(FF);
            double mu = 0, var = 0;
            for (int c = 0; c < C; ++c) { x[c] = xa[((size_t)im * L + l) * C + c]; mu += x[c]; }
            mu /= C;
            for (int c = 0; c < C; ++c) var += (x[c] - mu) * (x[c] - mu);
            double rs = 1.0 / std::sqrt(var / C + 1e-5);
            for (int c = 0; c < C; ++c) a[c] = h16((float)((x[c] - mu) * rs));
            for (int n = 0; n < C; ++n) { double acc = 0; for (int k = 0; k < C; ++k) acc += a[k] * h16(wq[(size_t)n * C + k]); qv[n] = acc > 0 ? acc + 1.0 : std::exp(acc); }
            for (int h = 0; h < NH; ++h) {
                double den = 1e-6;
                for (int d = 0; d < HD; ++d) den += qv[h * HD + d] * part[(size_t)im * KVS + NH * HD * HD + h * HD + d];
                for (int e = 0; e < HD; ++e) {
                    double acc = 0;
                    for (int d = 0; d < HD; ++d) acc += (double)h16((float)qv[h * HD + d]) * (double)h16(part[(size_t)im * KVS + h * HD * HD + d * HD + e]);
                    msg_[h * HD + e] = h16((float)(acc / den));
                }
            }
            for (int n = 0; n < C; ++n) { double acc = 0; for (int k = 0; k < C; ++k) acc += msg_[k] * h16(wm[(size_t)n * C + k]); x[n] += acc; }
            mu = 0; var = 0;
            for (int c = 0; c < C; ++c) mu += x[c];
            mu /= C;
            for (int c = 0; c < C; ++c) var += (x[c] - mu) * (x[c] - mu);
            rs = 1.0 / std::sqrt(var / C + 1e-5);
            for (int c = 0; c < C; ++c) a[c] = h16((float)((x[c] - mu) * rs));
            for (int n = 0; n < FF; ++n) { double acc = 0; for (int k = 0; k < C; ++k) acc += a[k] * h16(w1[(size_t)n * C + k]); hid[n] = h16((float)(0.5 * acc * (1.0 + std::erf(acc / std::sqrt(2.0))))); }
            for (int n = 0; n < C; ++n) { double acc = 0; for (int k = 0; k < FF; ++k) acc += hid[k] * h16(w2[(size_t)n * FF + k]); x[n] += acc; }
            for (int c = 0; c < C; ++c) {
                const float got = xt[(((size_t)0 * 64 + c / 4) * TILE + l) * 4 + c % 4];
                err = std::max(err, std::abs(got - x[c]));
                scale = std::max(scale, std::abs(x[c]));
            }
        }
        errs[3] = (float)(err / scale);
    }
    int flag = 0;
    ST(cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost));
    if (n_errs > 4) errs[4] = (float)flag;
    for (int i = 5; i < n_errs; ++i) errs[i] = 0.f;
#undef ST
    cudaFree(d_feat); cudaFree(d_wk); cudaFree(d_wv); cudaFree(d_wq); cudaFree(d_wm); cudaFree(d_w1); cudaFree(d_w2);
    cudaFree(d_ones); cudaFree(d_zeros); cudaFree(d_xt); cudaFree(d_post); cudaFree(d_part); cudaFree(d_img); cudaFree(d_flag);
    return 0;
}

}  // namespace oetr
