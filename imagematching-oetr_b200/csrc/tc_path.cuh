// Interface of the tcgen05 (OETR_PREC_FP16) path: fp16 UMMA operand images of the weights, its workspace
// slice, the encoder driver and the building-block self-test.  Implemented in tc_kernels.cu.
#pragma once
#include "oetr_common.cuh"

#include <vector>

namespace oetr {

struct TcWeights {
    __half* enc_img = nullptr;     // per encoder layer: the chunk stream the layer kernels consume (see tc_kernels.cu)
    __half* dec_img = nullptr;     // decoder cross-attention k/v projection chunk streams (2 layers)
    __half* head_img = nullptr;    // heatmap_conv.0: 9 tap GEMM images
    float* dec_t = nullptr;        // transposed fp32 decoder weights (k_decoder)
    float* dec_ts = nullptr;       // the same, sliced per cluster rank (k_decoder_cl)
    size_t enc_layer_halfs = 0, dec_layer_halfs = 0;
};

struct TcWorkspace {
    float* xt = nullptr;           // tile-blocked fp32 residual stream [tiles][64][128][4], per-image tiles
    float* xt_enc = nullptr;       // the same in the encoder's flat tiling (k_retile converts it into xt for the head)
    float* kv_part = nullptr;      // per-tile linear-attention partial summaries [tiles][KVS]
    float* dec_kvs = nullptr;      // [2 decoder layers][2B][KVS] cross-attention summaries
    __half* mimg = nullptr;        // [2B images] folded merge weights (hi/lo stage images, 256 KB each)
    float* ksum = nullptr;         // [2B][256] Ksum of every image for the current layer
    float* att = nullptr;          // [tiles][128] per-token <memory, hs> (head)
    float* gstat = nullptr;        // [tiles][32][2] per-tile GroupNorm partials (mean, M2)
    float* z = nullptr;            // [tiles][128] heat-map logits
    float* tlbr = nullptr;         // [2B][4]
    __half *qimg = nullptr, *kimg = nullptr, *vimg = nullptr, *oimg = nullptr;   // full attention: per-tile operand images
};

// CUDA-event bracket around every launch of the dominant kernel (k_enc with a query phase); read back by bench.py through
// oetr_profile_read.  Events are recorded on the launching stream, consecutive pairs = (begin, end).
struct KernelProfiler {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    size_t used = 0;
    void mark(cudaStream_t s) {
        if (!on) return;
        if (used == ev.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; }
            ev.push_back(e);
        }
        cudaEventRecord(ev[used++], s);
    }
};

void tc_carve(size_t& off, void* base, int B, int L1, int L2, TcWorkspace& w, bool full_attention);
int tc_prepare_weights(const float* d_w, const float* d_w9, const WLayout& L, TcWeights& out, char* msg, size_t msg_len);
void tc_free_weights(TcWeights& w);
// runs the 8 encoder layers and the decoder's cross-attention K/V summaries; writes token-major memory to X_out
// tile-blocked positional rows of one (hf,wf) geometry (cached by the handle, see oetr_abi.cu)
// OETR_TIMING=1: copies the device-side cycle accumulators (DBG_* in tc_tiles.cuh) to out; returns the slots copied
int tc_debug_read(unsigned long long* out, int n, int reset);
void tc_debug_enable(bool on);
// host-only consistency check of the encoder tiling; returns the number of flat tiles (> 0), -2 - tiles for the
// per-image tiling, or -1 with a message
int tc_check_geometry(int B, int L1, int L2, char* msg, size_t msg_len);
size_t tc_pos_tile_floats(int L);
void tc_pos_tiles(const float* d_pe, int max_w, int wf, int L, float* post, cudaStream_t s, LaunchCounter& lc);
int tc_encoder(const TcWeights& tw, const float* d_w, const float* h_w, const WLayout& L, const TcWorkspace& ws, const float* feat1,
               const float* feat2, int B, int hf1, int wf1, int hf2, int wf2, const float* post1, const float* post2,
               const float* mask1, const float* mask2,
               float* X_out, int* timeout_flag, KernelProfiler* prof, cudaStream_t s, LaunchCounter& lc, char* msg, size_t msg_len);
int tc_encoder_full(const TcWeights& tw, const float* d_w, const float* h_w, const WLayout& L, const TcWorkspace& ws, const float* feat1,
                    const float* feat2, int B, int hf1, int wf1, int hf2, int wf2, const float* post1, const float* post2,
                    float* X_out, int* timeout_flag, cudaStream_t s, LaunchCounter& lc, char* msg, size_t msg_len);
struct HeadGeom { int B, hf1, wf1, hf2, wf2, img_h1, img_w1, img_h2, img_w2, clamp; const float *mask1, *mask2; };
// fused fp32 query decoder (+ tlbr regression) -> hs_out [2B][256]; tcgen05 3x3 heat-map convolution -> Y scratch
// [B*L1+B*L2][256]; GroupNorm/ReLU/1x1 logits; softmax + soft-argmax + box assembly -> boxes1/boxes2 [B][4]
int tc_decoder_head(const TcWeights& tw, const float* d_w, const WLayout& L, const TcWorkspace& ws, const HeadGeom& hg,
                    float* hs_out, float* Y, float* boxes1, float* boxes2, float* dbg_cxy, float* dbg_tlbr,
                    int* timeout_flag, cudaStream_t s, LaunchCounter& lc, char* msg, size_t msg_len);
int tc_selftest(float* errs_host, int n_errs, char* msg, size_t msg_len);

}  // namespace oetr
