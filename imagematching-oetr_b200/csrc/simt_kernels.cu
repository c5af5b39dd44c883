// fp32 CUDA-core kernels of the OETR hot path (OETR_PREC_FP32 path, and the row-wise stages every path shares:
// layout change, LayerNorm+pos, decoder attention over one query, GroupNorm/softmax/soft-argmax head).
// Maths follows SURVEY.md Appendix A; reference lines are cited per kernel.
#include "oetr_common.cuh"

namespace oetr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------------------
// NCHW -> token-major (src/models/transformer.py:338-339: flatten(2).permute(0,2,1))
// ---------------------------------------------------------------------------------------------------------
__global__ void k_nchw_to_tokens(const float* __restrict__ feat, float* __restrict__ X, int L) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* src = feat + (size_t)b * C * L;
    float* dst = X + (size_t)b * L * C;
#pragma unroll
    for (int i = threadIdx.y; i < 32; i += 8) {
        int l = l0 + threadIdx.x;
        tile[i][threadIdx.x] = (l < L) ? src[(size_t)(c0 + i) * L + l] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = threadIdx.y; i < 32; i += 8) {
        int l = l0 + i;
        if (l < L) dst[(size_t)l * C + c0 + threadIdx.x] = tile[threadIdx.x][i];
    }
}
void nchw_to_tokens(const float* feat, float* X, int batch, int L, cudaStream_t s, LaunchCounter& lc) {
    dim3 grid((L + 31) / 32, C / 32, batch), block(32, 8);
    k_nchw_to_tokens<<<grid, block, 0, s>>>(feat, X, L);
    lc.n++;
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over 256 channels (+ positional term).  One warp per row, two-pass variance like ATen.
// src/models/transformer.py:118-126 (encoder), :236-245 (decoder)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_ln_pos(const float* __restrict__ in, const float* __restrict__ gamma,
                         const float* __restrict__ beta, const float* __restrict__ pos, int pos_rows,
                         float* __restrict__ out, int rows) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float4* src = reinterpret_cast<const float4*>(in + (size_t)warp * C);
    float4 a = src[lane], b = src[lane + 32];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (gamma != nullptr) {
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += v[i];
        const float mu = warp_sum(sum) * (1.f / C);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { float d = v[i] - mu; sq += d * d; }
        const float rstd = rsqrtf(warp_sum(sq) * (1.f / C) + LN_EPS);
        const float4 g0 = reinterpret_cast<const float4*>(gamma)[lane], g1 = reinterpret_cast<const float4*>(gamma)[lane + 32];
        const float4 b0 = reinterpret_cast<const float4*>(beta)[lane], b1 = reinterpret_cast<const float4*>(beta)[lane + 32];
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (v[i] - mu) * rstd * g[i] + bb[i];
    }
    if (pos != nullptr) {
        // pos_rows > 0: row r uses pos[r % pos_rows] (token position); pos_rows < 0: pos[r / -pos_rows] (stacked sets)
        const int prow = pos_rows > 0 ? warp % pos_rows : warp / (-pos_rows);
        const float4* p = reinterpret_cast<const float4*>(pos + (size_t)prow * C);
        float4 p0 = p[lane], p1 = p[lane + 32];
        v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w;
        v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
    }
    float4* dst = reinterpret_cast<float4*>(out + (size_t)warp * C);
    dst[lane] = make_float4(v[0], v[1], v[2], v[3]);
    dst[lane + 32] = make_float4(v[4], v[5], v[6], v[7]);
}
void ln_pos(const float* in, const float* gamma, const float* beta, const float* pos, int pos_rows,
            float* out, int rows, cudaStream_t s, LaunchCounter& lc) {
    if (rows <= 0) return;
    const int wpb = 8;
    k_ln_pos<<<(rows + wpb - 1) / wpb, wpb * 32, 0, s>>>(in, gamma, beta, pos, pos_rows, out, rows);
    lc.n++;
}

// ---------------------------------------------------------------------------------------------------------
// fp32 GEMM  C[M,N] = act(A[M,K] . W[N,K]^T + bias)   (nn.Linear layout; N % 64 == 0, K % 16 == 0)
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float x, int act) {
    switch (act) {
        case ACT_ELU1: return x > 0.f ? x + 1.f : expf(x);                  // elu(x)+1, linear_attention.py:12-13
        case ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));  // nn.GELU (erf form)
        case ACT_RELU: return fmaxf(x, 0.f);
        default: return x;
    }
}

__global__ void __launch_bounds__(256)
k_gemm_nt(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
          const float* __restrict__ bias, float* __restrict__ Cm, int ldc, int M, int N, int K,
          int act, int accumulate) {
    __shared__ __align__(16) float As[16][64 + 4];
    __shared__ __align__(16) float Bs[16][64 + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int lr = tid >> 2, lk = (tid & 3) * 4;        // loader: row 0..63, k offset 0,4,8,12
    const int ty = tid >> 4, tx = tid & 15;             // compute: 4 rows x 4 cols
    float acc[4][4] = {};
    const bool a_ok = (m0 + lr) < M;
    const float* a_ptr = A + (size_t)(m0 + lr) * lda + lk;
    const float* w_ptr = W + (size_t)(n0 + lr) * ldw + lk;
    for (int k0 = 0; k0 < K; k0 += 16) {
        float4 av = a_ok ? *reinterpret_cast<const float4*>(a_ptr + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 wv = *reinterpret_cast<const float4*>(w_ptr + k0);
        __syncthreads();
        As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
        Bs[lk + 0][lr] = wv.x; Bs[lk + 1][lr] = wv.y; Bs[lk + 2][lr] = wv.z; Bs[lk + 3][lr] = wv.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    const int n = n0 + tx * 4;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr) bv = *reinterpret_cast<const float4*>(bias + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float4* dst = reinterpret_cast<float4*>(Cm + (size_t)m * ldc + n);
        float4 r = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
        if (accumulate) {
            const float4 o = *dst;
            r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
        } else {
            r.x = act_apply(r.x, act); r.y = act_apply(r.y, act); r.z = act_apply(r.z, act); r.w = act_apply(r.w, act);
        }
        *dst = r;
    }
}
void gemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias, float* Cm, int ldc,
             int M, int N, int K, int act, int accumulate, cudaStream_t s, LaunchCounter& lc) {
    if (M <= 0) return;
    dim3 grid(N / 64, (M + 63) / 64);
    k_gemm_nt<<<grid, 256, 0, s>>>(A, lda, W, ldw, bias, Cm, ldc, M, N, K, act, accumulate);
    lc.n++;
}

// ---------------------------------------------------------------------------------------------------------
// Linear-attention summaries per (image, head): KV = Kf^T V, Ksum = sum_s Kf   (linear_attention.py:45-46)
// (the reference's v/S ... *S rescale, :43-48, is a numerical no-op and is dropped: SURVEY.md 8(a)-Q7)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_kv_reduce(const float* __restrict__ Kf, const float* __restrict__ V, float* __restrict__ kvs, int S) {
    __shared__ float Ks[32][33];
    __shared__ float Vs[32][33];
    const int h = blockIdx.x, b = blockIdx.y;
    const int tx = threadIdx.x, ty = threadIdx.y;            // tx = e (value channel), ty = d (key channel)
    const float* kb = Kf + (size_t)b * S * C + h * HD;
    const float* vb = V + (size_t)b * S * C + h * HD;
    float acc = 0.f, ksum = 0.f;
    for (int s0 = 0; s0 < S; s0 += 32) {
        const int srow = s0 + ty;
        Ks[ty][tx] = srow < S ? kb[(size_t)srow * C + tx] : 0.f;
        Vs[ty][tx] = srow < S ? vb[(size_t)srow * C + tx] : 0.f;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float kd = Ks[j][ty];
            acc = fmaf(kd, Vs[j][tx], acc);
            ksum += kd;
        }
        __syncthreads();
    }
    float* o = kvs + (size_t)b * KVS;
    o[h * HD * HD + ty * HD + tx] = acc;
    if (tx == 0) o[NH * HD * HD + h * HD + ty] = ksum;
}
void kv_reduce(const float* Kf, const float* V, float* kvs, int batch, int S, cudaStream_t s, LaunchCounter& lc) {
    if (batch <= 0) return;
    k_kv_reduce<<<dim3(NH, batch), dim3(32, 32), 0, s>>>(Kf, V, kvs, S);
    lc.n++;
}

// out = Qf . KV / (Qf . Ksum + eps)   (linear_attention.py:46-48)
__global__ void __launch_bounds__(256)
k_linattn_apply(const float* __restrict__ Qf, const float* __restrict__ kvs, float* __restrict__ O,
                int rows_per_image) {
    __shared__ float qs[C];
    const int r = blockIdx.x;
    const int c = threadIdx.x, h = c >> 5, e = c & 31;
    const float* kv = kvs + (size_t)(r / rows_per_image) * KVS;
    const float q = Qf[(size_t)r * C + c];
    qs[c] = q;
    const float den = warp_sum(q * kv[NH * HD * HD + c]) + ATTN_EPS;      // lane e doubles as d here
    __syncthreads();
    const float* kvh = kv + h * HD * HD + e;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc = fmaf(qs[h * HD + d], kvh[d * HD], acc);
    O[(size_t)r * C + c] = acc / den;
}
void linattn_apply(const float* Qf, const float* kvs, float* O, int rows, int rows_per_image,
                   cudaStream_t s, LaunchCounter& lc) {
    if (rows <= 0) return;
    k_linattn_apply<<<rows, 256, 0, s>>>(Qf, kvs, O, rows_per_image);
    lc.n++;
}

// ---------------------------------------------------------------------------------------------------------
// Softmax attention (linear_attention.py:59-87): one warp per (image, head, query), keys staged per CTA.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_full_attention(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                 float* __restrict__ O, int L, int S) {
    __shared__ float Ks[32][33];
    __shared__ float Vs[32][33];
    __shared__ float Qs[8][32];
    const int b = blockIdx.z, h = blockIdx.y;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = blockIdx.x * 8 + w;
    const bool active = l < L;
    Qs[w][lane] = active ? q[((size_t)b * L + l) * C + h * HD + lane] : 0.f;
    const float scale = 0.17677669529663688110f;   // 1/sqrt(32)
    float m = -INFINITY, den = 0.f, o = 0.f;
    const float* kb = k + (size_t)b * S * C + h * HD;
    const float* vb = v + (size_t)b * S * C + h * HD;
    for (int s0 = 0; s0 < S; s0 += 32) {
        __syncthreads();
        for (int i = w; i < 32; i += 8) {
            const int srow = s0 + i;
            Ks[i][lane] = srow < S ? kb[(size_t)srow * C + lane] : 0.f;
            Vs[i][lane] = srow < S ? vb[(size_t)srow * C + lane] : 0.f;
        }
        __syncthreads();
        float sc = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) sc = fmaf(Qs[w][d], Ks[lane][d], sc);
        sc = (s0 + lane < S) ? sc * scale : -INFINITY;
        const float m_new = fmaxf(m, warp_max(sc));
        const float p = expf(sc - m_new);
        const float corr = expf(m - m_new);
        den = den * corr + warp_sum(p);
        o *= corr;
#pragma unroll
        for (int j = 0; j < 32; ++j) o = fmaf(__shfl_sync(0xffffffffu, p, j), Vs[j][lane], o);
        m = m_new;
    }
    if (active) O[((size_t)b * L + l) * C + h * HD + lane] = o / den;
}
void full_attention(const float* q, const float* k, const float* v, float* O, int batch, int L, int S,
                    cudaStream_t s, LaunchCounter& lc) {
    if (batch <= 0) return;
    k_full_attention<<<dim3((L + 7) / 8, NH, batch), 256, 0, s>>>(q, k, v, O, L, S);
    lc.n++;
}

// ---------------------------------------------------------------------------------------------------------
// Head (src/model.py:145-191)
// ---------------------------------------------------------------------------------------------------------
// heat = memory * <memory, hs>   (:147-155)
__global__ void k_heat_scale(const float* __restrict__ M, const float* __restrict__ hs, float* __restrict__ G,
                             int rows, int L) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float4* m = reinterpret_cast<const float4*>(M + (size_t)warp * C);
    const float4* hv = reinterpret_cast<const float4*>(hs + (size_t)(warp / L) * C);
    const float4 a = m[lane], b = m[lane + 32], x = hv[lane], y = hv[lane + 32];
    const float att = warp_sum(a.x * x.x + a.y * x.y + a.z * x.z + a.w * x.w +
                               b.x * y.x + b.y * y.y + b.z * y.z + b.w * y.w);
    float4* g = reinterpret_cast<float4*>(G + (size_t)warp * C);
    g[lane] = make_float4(a.x * att, a.y * att, a.z * att, a.w * att);
    g[lane + 32] = make_float4(b.x * att, b.y * att, b.z * att, b.w * att);
}
void heat_scale(const float* M, const float* hs, float* G, int rows, int L, cudaStream_t s, LaunchCounter& lc) {
    if (rows <= 0) return;
    k_heat_scale<<<(rows + 7) / 8, 256, 0, s>>>(M, hs, G, rows, L);
    lc.n++;
}

// zero-padded spatial shift of a token-major map: one tap of the 3x3 convolution's im2col
__global__ void k_shift_tokens(const float* __restrict__ in, float* __restrict__ out, int batch, int hf, int wf,
                               int dy, int dx) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // float4 index
    const size_t total = (size_t)batch * hf * wf * (C / 4);
    if (idx >= total) return;
    const int c4 = idx % (C / 4);
    size_t t = idx / (C / 4);
    const int x = t % wf; t /= wf;
    const int y = t % hf;
    const int b = t / hf;
    const int ys = y + dy, xs = x + dx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ys >= 0 && ys < hf && xs >= 0 && xs < wf)
        v = reinterpret_cast<const float4*>(in)[(((size_t)b * hf + ys) * wf + xs) * (C / 4) + c4];
    reinterpret_cast<float4*>(out)[idx] = v;
}
void shift_tokens(const float* in, float* out, int batch, int hf, int wf, int dy, int dx,
                  cudaStream_t s, LaunchCounter& lc) {
    const size_t total = (size_t)batch * hf * wf * (C / 4);
    if (total == 0) return;
    k_shift_tokens<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, out, batch, hf, wf, dy, dx);
    lc.n++;
}

__device__ float block_sum(float v, float* red) {       // 256 threads
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += red[i];
    return r;
}
__device__ float block_max(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = fmaxf(r, red[i]);
    return r;
}

// GroupNorm(32) -> ReLU -> 1x1 conv -> softmax over tokens -> soft-argmax (:65-77,:157-184);
// tlbr = sigmoid(W2 relu(W0 hs) + b2) (:59-63,:188-191); box assembly (models/utils.py:16-28 / model.py:193-211)
__global__ void __launch_bounds__(256)
k_head_finalize(HeadParams p) {
    extern __shared__ float z[];                 // [L] logits
    __shared__ float red[8];
    __shared__ float g_mean[32], g_rstd[32];
    __shared__ float hvec[C], hid[C];
    __shared__ float tl[4];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int L = p.hf * p.wf;
    const float* Y = p.Y + (size_t)b * L * C;
    // ---- GroupNorm statistics: thread = channel; group = 8 adjacent channels = 8 adjacent lanes
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += Y[(size_t)l * C + tid];
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float mean = s / (8.f * L);
    float q = 0.f;
    for (int l = 0; l < L; ++l) { const float d = Y[(size_t)l * C + tid] - mean; q = fmaf(d, d, q); }
    q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
    if ((tid & 7) == 0) { g_mean[tid >> 3] = mean; g_rstd[tid >> 3] = rsqrtf(q / (8.f * L) + GN_EPS); }
    hvec[tid] = p.hs[(size_t)b * C + tid];
    __syncthreads();
    // ---- logits: warp per token, lane covers channels lane*4.. and 128+lane*4..
    float sc[8], sh[8], w3[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = (i < 4) ? lane * 4 + i : 128 + lane * 4 + (i - 4);
        const float rs = g_rstd[c >> 3] * p.gn_g[c];
        sc[i] = rs; sh[i] = p.gn_b[c] - g_mean[c >> 3] * rs; w3[i] = p.w3[c];
    }
    const float b3 = p.b3[0];
    for (int l = w; l < L; l += 8) {
        const float4 a = reinterpret_cast<const float4*>(Y + (size_t)l * C)[lane];
        const float4 c4 = reinterpret_cast<const float4*>(Y + (size_t)l * C)[lane + 32];
        const float v[8] = {a.x, a.y, a.z, a.w, c4.x, c4.y, c4.z, c4.w};
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(w3[i], fmaxf(fmaf(v[i], sc[i], sh[i]), 0.f), acc);
        acc = warp_sum(acc);
        if (lane == 0) z[l] = (p.mask && p.mask[(size_t)b * L + l] == 0.f) ? -1e9f : acc + b3;
    }
    __syncthreads();
    // ---- softmax + soft-argmax over the (x+0.5, y+0.5)*stride grid; stride = img_h / hf for both axes (:176-181)
    float mx = -INFINITY;
    for (int l = tid; l < L; l += 256) mx = fmaxf(mx, z[l]);
    mx = block_max(mx, red);
    const float stride = (float)(p.img_h / p.hf);
    float se = 0.f, sx = 0.f, sy = 0.f;
    for (int l = tid; l < L; l += 256) {
        const float e = expf(z[l] - mx);
        se += e;
        sx = fmaf(e, ((float)(l % p.wf) + 0.5f) * stride, sx);
        sy = fmaf(e, ((float)(l / p.wf) + 0.5f) * stride, sy);
    }
    se = block_sum(se, red); sx = block_sum(sx, red); sy = block_sum(sy, red);
    const float cx = sx / se, cy = sy / se;
    // ---- tlbr regression
    for (int r = w; r < C; r += 8) {
        const float* wr = p.tl_w0 + (size_t)r * C;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(wr[lane + 32 * i], hvec[lane + 32 * i], acc);
        acc = warp_sum(acc);
        if (lane == 0) hid[r] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    if (w < 4) {
        const float* wr = p.tl_w2 + (size_t)w * C;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(wr[lane + 32 * i], hid[lane + 32 * i], acc);
        acc = warp_sum(acc);
        if (lane == 0) tl[w] = 1.f / (1.f + expf(-(acc + p.tl_b2[w])));
    }
    __syncthreads();
    if (tid == 0) {
        const float W_ = (float)p.img_w, H_ = (float)p.img_h;
        float x1 = cx - tl[1] * W_, y1 = cy - tl[0] * H_, x2 = cx + tl[3] * W_, y2 = cy + tl[2] * H_;
        if (p.clamp) {
            x1 = fminf(fmaxf(x1, 0.f), W_); x2 = fminf(fmaxf(x2, 0.f), W_);
            y1 = fminf(fmaxf(y1, 0.f), H_); y2 = fminf(fmaxf(y2, 0.f), H_);
        }
        float* o = p.boxes + (size_t)b * 4;
        o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
        if (p.dbg_cxy) { p.dbg_cxy[b * 2] = cx; p.dbg_cxy[b * 2 + 1] = cy; }
        if (p.dbg_tlbr) { for (int i = 0; i < 4; ++i) p.dbg_tlbr[b * 4 + i] = tl[i]; }
    }
}
__global__ void k_row_scale(float* __restrict__ X, const float* __restrict__ mask, int rows) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // one float4 per thread
    if (idx >= rows * (C / 4)) return;
    const float m = mask[idx / (C / 4)];
    float4 v = reinterpret_cast<float4*>(X)[idx];
    v.x *= m; v.y *= m; v.z *= m; v.w *= m;
    reinterpret_cast<float4*>(X)[idx] = v;
}
void row_scale(float* X, const float* mask, int rows, cudaStream_t s, LaunchCounter& lc) {
    if (rows <= 0 || mask == nullptr) return;
    k_row_scale<<<(rows * (C / 4) + 255) / 256, 256, 0, s>>>(X, mask, rows);
    lc.n++;
}

void head_finalize(const HeadParams& p, cudaStream_t s, LaunchCounter& lc) {
    if (p.batch <= 0) return;
    const size_t smem = (size_t)p.hf * p.wf * sizeof(float);
    k_head_finalize<<<p.batch, 256, smem, s>>>(p);
    lc.n++;
}

}  // namespace oetr
