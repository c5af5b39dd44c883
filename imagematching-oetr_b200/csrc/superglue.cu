// SuperGlue's two hot operators on the device, the heaviest consumer downstream of the overlap
// boxes in the reference's pipeline (evaluation.py:125-224 -> third_party/SuperGluePretrainedNetwork/models/superglue.py):
//   k_sg_attention   `attention(query, key, value)` (superglue.py:86-90) for the 4 heads x 64 dims of MultiHeadedAttention
//                    (:93-108): softmax(Q K^T / 8) V with an online softmax over 64-key tiles -- the [N, M] score matrix
//                    never exists in memory (the reference materialises it twice per layer, 18 layers).  fp32 on the CUDA
//                    cores: exact-parity arithmetic (SuperGlue's match decisions are mutual-argmax + a threshold on
//                    exp(score)); the contraction sizes are tiny next to OETR's (<= 38 GFLOP for 2048 keypoints over all
//                    layers), a tcgen05 version with split operands is listed as next work in DESIGN.md.
//   k_sg_rows / k_sg_cols / k_sg_finish   `log_optimal_transport` (superglue.py:150-184): log-space Sinkhorn iterations on
//                    the score matrix augmented by the dustbin row / column.  The augmented matrix is never built (the
//                    dustbin entries are the scalar alpha), u and v live in a small workspace, every half-iteration is one
//                    pass over the scores with an online log-sum-exp (rows: a warp per row; columns: 32 columns per block,
//                    rows strided over 8 warps), all launches stream-ordered.
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
__device__ __forceinline__ void lse_add(LSE& a, float x) {
    if (x > a.m) { a.s = a.s * expf(a.m - x) + 1.f; a.m = x; }
    else a.s += expf(x - a.m);
}
__device__ __forceinline__ void lse_merge(LSE& a, const LSE& b) {
    if (b.m == -INFINITY) return;
    if (b.m > a.m) { a.s = a.s * expf(a.m - b.m) + b.s; a.m = b.m; }
    else a.s += b.s * expf(b.m - a.m);
}

// u[i] = log_mu[i] - logsumexp_j (Z[i][j] + v[j]), i in [0, m], j in [0, n]; Z = scores inside, alpha in the dustbins
__global__ void __launch_bounds__(256) k_sg_rows(const float* __restrict__ scores, float alpha, const float* __restrict__ vv,
                                                 float* __restrict__ u, int m, int n, float norm, int first) {
    const int b = blockIdx.y, row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row > m) return;
    const float* sc = scores + (size_t)b * m * n + (size_t)row * n;
    const float* v = vv + (size_t)b * (n + 1);
    LSE a{-INFINITY, 0.f};
    for (int j = lane; j <= n; j += 32) {
        const float z = (row < m && j < n) ? sc[j] : alpha;
        lse_add(a, z + (first ? 0.f : v[j]));
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        LSE o{__shfl_xor_sync(0xffffffffu, a.m, w), __shfl_xor_sync(0xffffffffu, a.s, w)};
        lse_merge(a, o);
    }
    if (lane == 0) u[(size_t)b * (m + 1) + row] = (row < m ? norm : logf((float)n) + norm) - (a.m + logf(a.s));
}
// v[j] = log_nu[j] - logsumexp_i (Z[i][j] + u[i])
__global__ void __launch_bounds__(256) k_sg_cols(const float* __restrict__ scores, float alpha, const float* __restrict__ uu,
                                                 float* __restrict__ v, int m, int n, float norm) {
    __shared__ float sm_m[8][33], sm_s[8][33];
    const int b = blockIdx.y, lane = threadIdx.x & 31, ty = threadIdx.x >> 5, col = blockIdx.x * 32 + lane;
    const float* sc = scores + (size_t)b * m * n;
    const float* u = uu + (size_t)b * (m + 1);
    LSE a{-INFINITY, 0.f};
    if (col <= n)
        for (int i = ty; i <= m; i += 8) {
            const float z = (i < m && col < n) ? sc[(size_t)i * n + col] : alpha;
            lse_add(a, z + u[i]);
        }
    sm_m[ty][lane] = a.m; sm_s[ty][lane] = a.s;
    __syncthreads();
    if (ty == 0 && col <= n) {
        for (int t = 1; t < 8; ++t) { LSE o{sm_m[t][lane], sm_s[t][lane]}; lse_merge(a, o); }
        v[(size_t)b * (n + 1) + col] = (col < n ? norm : logf((float)m) + norm) - (a.m + logf(a.s));
    }
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

size_t oetr_sg_transport_workspace_bytes(int batch, int m, int n) {
    return (size_t)(batch > 0 ? batch : 0) * (size_t)(m + n + 2) * sizeof(float);
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
    const float norm = -logf((float)(m + n));
    if (iters == 0) {
        const cudaError_t e0 = cudaMemsetAsync(workspace, 0, oetr_sg_transport_workspace_bytes(batch, m, n), s);
        if (e0 != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_optimal_transport: %s", cudaGetErrorString(e0));
    }
    for (int it = 0; it < iters; ++it) {
        k_sg_rows<<<dim3((m + 1 + 7) / 8, batch), 256, 0, s>>>(scores, alpha, v, u, m, n, norm, it == 0);
        k_sg_cols<<<dim3((n + 1 + 31) / 32, batch), 256, 0, s>>>(scores, alpha, u, v, m, n, norm);
    }
    k_sg_finish<<<dim3((n + 1 + 255) / 256, m + 1, batch), 256, 0, s>>>(scores, alpha, u, v, out, m, n, norm);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sfail(OETR_E_CUDA, "oetr_sg_optimal_transport: %s", cudaGetErrorString(e));
    return OETR_OK;
}

}  // extern "C"
