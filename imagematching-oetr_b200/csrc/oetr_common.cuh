// Shared definitions of the OETR hot-path library: model constants, the canonical packed-weight layout
// (documented in include/oetr_b200.h) and the launch-wrapper prototypes of the two kernel families.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stddef.h>
#include <stdint.h>

namespace oetr {

constexpr int C = 256;        // d_model (src/model.py:44: last_layer // 4)
constexpr int NH = 8;         // heads (src/model.py:82-84)
constexpr int HD = 32;        // head dim
constexpr int FF = 512;       // MLP hidden (src/models/transformer.py:91-95)
constexpr int N_ENC = 8;      // ['self','cross'] * 4
constexpr int N_DEC = 2;
constexpr int KVS = NH * HD * HD + NH * HD;   // floats of one linear-attention summary: KV[8][32][32] | Ksum[8][32]
constexpr float LN_EPS = 1e-5f;
constexpr float ATTN_EPS = 1e-6f;
constexpr float GN_EPS = 1e-5f;

// ---- canonical packed-weight layout (offsets in floats) ------------------------------------------------
struct EncW { size_t wq, wk, wv, wm, w1, w2, lnq_g, lnq_b, lnkv_g, lnkv_b, ln2_g, ln2_b; };
struct AttW { size_t wq, bq, wk, bk, wv, bv, wm; };
struct DecW { AttW sa, ca; size_t w1, w2, ln1_g, ln1_b, ln2_g, ln2_b, ln3_g, ln3_b; };
struct WLayout {
    EncW enc[N_ENC];
    DecW dec[N_DEC];
    size_t qe1, qe2;
    size_t tl_w0, tl_w2, tl_b2;
    size_t hm_w0, hm_b0, hm_gn_g, hm_gn_b, hm_w3, hm_b3;
    size_t total;
};

inline WLayout make_layout() {
    WLayout L{};
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += n; return r; };
    for (int i = 0; i < N_ENC; ++i) {
        EncW& e = L.enc[i];
        e.wq = take(C * C); e.wk = take(C * C); e.wv = take(C * C); e.wm = take(C * C);
        e.w1 = take(FF * C); e.w2 = take(C * FF);
        e.lnq_g = take(C); e.lnq_b = take(C); e.lnkv_g = take(C); e.lnkv_b = take(C);
        e.ln2_g = take(C); e.ln2_b = take(C);
    }
    for (int j = 0; j < N_DEC; ++j) {
        DecW& d = L.dec[j];
        for (AttW* a : {&d.sa, &d.ca}) {
            a->wq = take(C * C); a->bq = take(C); a->wk = take(C * C); a->bk = take(C);
            a->wv = take(C * C); a->bv = take(C); a->wm = take(C * C);
        }
        d.w1 = take(FF * C); d.w2 = take(C * FF);
        d.ln1_g = take(C); d.ln1_b = take(C); d.ln2_g = take(C); d.ln2_b = take(C);
        d.ln3_g = take(C); d.ln3_b = take(C);
    }
    L.qe1 = take(C); L.qe2 = take(C);
    L.tl_w0 = take(C * C); L.tl_w2 = take(4 * C); L.tl_b2 = take(4);
    L.hm_w0 = take((size_t)C * C * 9); L.hm_b0 = take(C); L.hm_gn_g = take(C); L.hm_gn_b = take(C);
    L.hm_w3 = take(C); L.hm_b3 = take(1);
    L.total = o;
    return L;
}

// activation selectors of the fp32 GEMM epilogue
enum Act { ACT_NONE = 0, ACT_ELU1 = 1, ACT_GELU = 2, ACT_RELU = 3 };

struct LaunchCounter { int n = 0; };

// ---- fp32 CUDA-core kernels (simt_kernels.cu) -----------------------------------------------------------
// X[b][l][c] = feat[b][c][l]
void nchw_to_tokens(const float* feat, float* X, int batch, int L, cudaStream_t s, LaunchCounter& lc);
// out[r][:] = (gamma ? LN(in[r][:]) : in[r][:]) + (pos ? pos[pos_rows > 0 ? r % pos_rows : r / -pos_rows][:] : 0)
void ln_pos(const float* in, const float* gamma, const float* beta, const float* pos, int pos_rows,
            float* out, int rows, cudaStream_t s, LaunchCounter& lc);
// Cmat[M,N] (ldc) = act(A[M,K](lda) . W[N,K]^T + bias) ; accumulate: Cmat += (no act)
void gemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias, float* Cmat, int ldc,
             int M, int N, int K, int act, int accumulate, cudaStream_t s, LaunchCounter& lc);
// per image b, head h: KV[d][e] = sum_s Kf[b,s,h,d] * V[b,s,h,e] ; Ksum[d] = sum_s Kf[b,s,h,d]   (Kf already elu+1)
void kv_reduce(const float* Kf, const float* V, float* kvs, int batch, int S, cudaStream_t s, LaunchCounter& lc);
// O[r][h*32+e] = sum_d Qf[r][h*32+d] KV[img][h][d][e] / (Qf[r][h,:].Ksum[img][h,:] + eps), img = r / rows_per_image
void linattn_apply(const float* Qf, const float* kvs, float* O, int rows, int rows_per_image,
                   cudaStream_t s, LaunchCounter& lc);
// softmax(q k^T / sqrt(32)) v per head; q [batch,L,256], k,v [batch,S,256]
void full_attention(const float* q, const float* k, const float* v, float* O, int batch, int L, int S,
                    cudaStream_t s, LaunchCounter& lc);
// G[r][:] = M[r][:] * <M[r][:], hs[r / L][:]>
void heat_scale(const float* M, const float* hs, float* G, int rows, int L, cudaStream_t s, LaunchCounter& lc);
// out[b][y][x][:] = in[b][y+dy][x+dx][:] (zero outside)
void shift_tokens(const float* in, float* out, int batch, int hf, int wf, int dy, int dx,
                  cudaStream_t s, LaunchCounter& lc);
struct HeadParams {
    const float* Y;        // conv3x3 output incl. bias, token-major [batch*L][256]
    const float* hs;       // [batch][256]
    const float *gn_g, *gn_b, *w3, *b3, *tl_w0, *tl_w2, *tl_b2;
    int batch, hf, wf, img_h, img_w, clamp;
    float *boxes, *dbg_cxy, *dbg_tlbr;   // dbg nullable
    const float* mask;     // nullable [batch][hf*wf]: logits where mask == 0 are filled with -1e9 (src/model.py:167-171)
};
// GroupNorm(32)+ReLU+1x1 conv+softmax+soft-argmax, tlbr MLP+sigmoid, box assembly; one CTA per image
void head_finalize(const HeadParams& p, cudaStream_t s, LaunchCounter& lc);
// X[r][:] *= mask[r] for r < rows (256 channels): the q / kv masks of LinearAttention (linear_attention.py:36-41)
void row_scale(float* X, const float* mask, int rows, cudaStream_t s, LaunchCounter& lc);

}  // namespace oetr
