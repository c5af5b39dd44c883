// SuperGlue's two hot operators on the device, the heaviest consumer downstream of the overlap
// boxes in the reference's pipeline (evaluation.py:125-224 -> third_party/SuperGluePretrainedNetwork/models/superglue.py):
//   k_sg_attention   `attention(query, key, value)` (superglue.py:86-90) for the 4 heads x 64 dims of MultiHeadedAttention
//                    (:93-108): softmax(Q K^T / 8) V with an online softmax over 64-key tiles -- the [N, M] score matrix
//                    never exists in memory (the reference materialises it twice per layer, 18 layers).  fp32 on the CUDA
//                    cores: exact-parity arithmetic (SuperGlue's match decisions are mutual-argmax + a threshold on
//                    exp(score)); the contraction sizes are tiny next to OETR's (<= 38 GFLOP for 2048 keypoints over all
//                    layers), a tcgen05 version with split operands is listed as next work in DESIGN.md.
//   k_sg_half / k_sg_transpose / k_sg_finish   `log_optimal_transport` (superglue.py:150-184): log-space Sinkhorn iterations
//                    on the score matrix augmented by the dustbin row / column.  The augmented matrix is never built (the
//                    dustbin entries are the scalar alpha), u and v live in a small workspace, every half-iteration is one
//                    pass over the L2-resident scores with a chunked log-sum-exp (a warp per row, 8 independent loads and
//                    exponentials per lane and step); the column update runs the same kernel on a transposed copy made
//                    once per call, so both passes are coalesced.  All launches stream-ordered.
// Layout: channel-major like the reference's Conv1d tensors, q [batch][256][N] with channel c = d * 4 + h
// (`.view(batch, dim, heads, -1)`, superglue.py:103-104).
#include "../../include/oetr_b200.h"

#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
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

// ---- log-space optimal transport ------------------------------------------------------------------------------------
struct LSE { float m, s; };
__device__ __forceinline__ void lse_merge(LSE& a, const LSE& b) {
    const float m = fmaxf(a.m, b.m);
    if (m == -INFINITY) return;
    a.s = a.s * expf(a.m - m) + b.s * expf(b.m - m);
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
    for (int i = 0; i < 8; ++i) sum += expf(x[i] - m);
    a.s = a.s * expf(a.m - m) + sum;
    a.m = m;
}

// One half-iteration: out[r] = log_marg[r] - logsumexp_c (Z[r][c] + in[c]) for the R+1 rows of the augmented matrix whose
// inner part `mat` is [R][Cn] row-major (the scores for the u-update, their transpose for the v-update) and whose dustbin
// row / column is the scalar alpha.  A warp per row: coalesced loads, 8 independent loads in flight per lane.
__global__ void __launch_bounds__(256) k_sg_half(const float* __restrict__ mat, float alpha, const float* __restrict__ in,
                                                 float* __restrict__ out, int R, int Cn, float norm, float log_other, int first) {
    const int b = blockIdx.y, row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row > R) return;
    const float* mr = mat + (size_t)b * R * Cn + (size_t)row * Cn;
    const float* iv = in + (size_t)b * (Cn + 1);
    const bool inner = row < R;
    LSE a{-INFINITY, 0.f};
    for (int c0 = 0; c0 <= Cn; c0 += 256) {
        float x[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int c = c0 + t * 32 + lane;
            float z = -INFINITY;
            if (c <= Cn) z = ((inner && c < Cn) ? __ldg(mr + c) : alpha) + (first ? 0.f : iv[c]);
            x[t] = z;
        }
        lse_chunk(a, x);
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        LSE o{__shfl_xor_sync(0xffffffffu, a.m, w), __shfl_xor_sync(0xffffffffu, a.s, w)};
        lse_merge(a, o);
    }
    if (lane == 0) out[(size_t)b * (R + 1) + row] = (inner ? norm : log_other + norm) - (a.m + logf(a.s));
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
// out[i][j] = Z[i][j] + u[i] + v[j] - norm        ([m+1][n+1], the matrix the reference returns)
__global__ void __launch_bounds__(256) k_sg_finish(const float* __restrict__ scores, float alpha, const float* __restrict__ uu,
                                                   const float* __restrict__ vv, float* __restrict__ out, int m, int n, float norm) {
    const int b = blockIdx.z, j = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
    if (j > n) return;
    const float z = (i < m && j < n) ? scores[(size_t)b * m * n + (size_t)i * n + j] : alpha;
    out[(size_t)b * (m + 1) * (n + 1) + (size_t)i * (n + 1) + j] = z + uu[(size_t)b * (m + 1) + i] + vv[(size_t)b * (n + 1) + j] - norm;
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

int oetr_sg_attention(const float* query, const float* key, const float* value, float* out, int batch, int n, int m, void* stream) {
    if (!query || !key || !value || !out) return sfail(OETR_E_ARG, "oetr_sg_attention: null argument");
    if (batch < 1 || n < 1 || m < 1 || batch > 65535) return sfail(OETR_E_SHAPE, "oetr_sg_attention: batch %d, %d queries, %d keys", batch, n, m);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_sg_mu);
        if (dev < 64 && !g_sg_attr[dev]) {
            e = cudaFuncSetAttribute(k_sg_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM);
            if (e == cudaSuccess) g_sg_attr[dev] = true;
        }
    }
    if (e != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_attention: %s", cudaGetErrorString(e));
    k_sg_attention<<<dim3((n + BQ - 1) / BQ, SG_H, batch), 256, ATT_SMEM, static_cast<cudaStream_t>(stream)>>>(query, key, value, out, n, m);
    e = cudaGetLastError();
    if (e != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_attention: %s", cudaGetErrorString(e));
    return OETR_OK;
}

// u [batch][m+1] | v [batch][n+1] | transposed scores [batch][n][m]
size_t oetr_sg_transport_workspace_bytes(int batch, int m, int n) {
    if (batch < 1 || m < 1 || n < 1) return 0;
    return (size_t)batch * ((size_t)(m + n + 2) + (size_t)m * n) * sizeof(float);
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
    float* u = static_cast<float*>(workspace);
    float* v = u + (size_t)batch * (m + 1);
    float* st = v + (size_t)batch * (n + 1);
    const float norm = -logf((float)(m + n));
    if (iters == 0) {
        const cudaError_t e0 = cudaMemsetAsync(workspace, 0, (size_t)batch * (m + n + 2) * sizeof(float), s);
        if (e0 != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_optimal_transport: %s", cudaGetErrorString(e0));
    } else {
        k_sg_transpose<<<dim3((n + 31) / 32, (m + 31) / 32, batch), 256, 0, s>>>(scores, st, m, n);
    }
    for (int it = 0; it < iters; ++it) {
        k_sg_half<<<dim3((m + 1 + 7) / 8, batch), 256, 0, s>>>(scores, alpha, v, u, m, n, norm, logf((float)n), it == 0);
        k_sg_half<<<dim3((n + 1 + 7) / 8, batch), 256, 0, s>>>(st, alpha, u, v, n, m, norm, logf((float)m), 0);
    }
    k_sg_finish<<<dim3((n + 1 + 255) / 256, m + 1, batch), 256, 0, s>>>(scores, alpha, u, v, out, m, n, norm);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_optimal_transport: %s", cudaGetErrorString(e));
    return OETR_OK;
}

}  // extern "C"
