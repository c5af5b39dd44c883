/*
 * oetr_b200.h -- C ABI of the B200-native OETR hot path (liboetr_b200.so).
 *
 * The reference (TencentYoutuResearch/ImageMatching-OETR) is pure Python/PyTorch and has no FFI; this header is
 * the drop-in boundary a maintainer binds with ctypes.  Each entry point names the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross this boundary.
 *   - every function returns an int status (0 = OETR_OK, <0 = error); oetr_last_error() gives the message of
 *     the last failing call on the calling thread.  No exception ever crosses the boundary.
 *   - all device pointers are caller-owned (e.g. torch allocations); the handle owns only its private copy of
 *     the weights (and the constant position-encoding table derived at create time).
 *   - oetr_forward is asynchronous and stream-ordered: no hidden synchronisation, no allocation.
 *   - the library targets sm_100a only; on any other device oetr_create fails with OETR_E_ARCH (there is no
 *     fallback path, CPU or otherwise).
 */
#ifndef OETR_B200_H_
#define OETR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OETR_API __attribute__((visibility("default")))
#else
#define OETR_API
#endif

#define OETR_ABI_VERSION 1

/* status codes */
#define OETR_OK        0
#define OETR_E_ARG    (-1)   /* null pointer / bad enum / bad size                                         */
#define OETR_E_SHAPE  (-2)   /* feature-map shape outside what the reference supports (1..100 per side)    */
#define OETR_E_ARCH   (-3)   /* current device is not sm_100                                               */
#define OETR_E_CUDA   (-4)   /* a CUDA runtime call failed (message in oetr_last_error)                    */
#define OETR_E_NOMEM  (-5)   /* device allocation failed / workspace too small                             */

/* attention_mode: which attention the encoder layers use.
 *   LINEAR = src/models/linear_attention.py:16-50 (what OETR ships: src/model.py:82-84 never passes a mode)
 *   FULL   = src/models/linear_attention.py:53-87 (QueryTransformer(attention_mode='full'), encoder only) */
#define OETR_ATTN_LINEAR 0
#define OETR_ATTN_FULL   1

/* operand_precision: arithmetic of the dense contractions.
 *   FP32 = CUDA-core fp32 FMA everywhere (reference-grade, slow)
 *   FP16 = tcgen05 tensor-core MMA, fp16 operands / fp32 TMEM accumulation for the encoder GEMMs and the
 *          linear-attention contractions; everything row-wise (LayerNorm, elu, GELU, softmax, GroupNorm)
 *          stays fp32.  (plain bf16 operands miss the 1e-3 parity bar, SURVEY.md finding 3.) */
#define OETR_PREC_FP32 0
#define OETR_PREC_FP16 1

typedef struct oetr_handle oetr_handle;

/* Number of fp32 values in the packed weight blob, in the canonical order below.  (6 443 525)            */
OETR_API size_t oetr_packed_weight_count(void);

/*
 * Create a handle from the packed hot-path weights (fp32, host OR device pointer, copied).
 * Replaces: OETR.__init__'s construction of QueryTransformer / tlbr_reg / heatmap_conv / query_embed* /
 * PositionEncodingSine (src/model.py:58-86) + load_state_dict (dloc/core/overlaps/oetr.py:36-42).
 *
 * Canonical order of `weights` (each tensor row-major with the shape torch's state_dict gives it):
 *   for i in 0..7,  prefix "transformer.encoder.<i>.":
 *       q_proj.weight[256,256] k_proj.weight[256,256] v_proj.weight[256,256] merge.weight[256,256]
 *       mlp.0.weight[512,256] mlp.2.weight[256,512]
 *       pre_norm_q.weight[256] pre_norm_q.bias[256] pre_norm_kv.weight[256] pre_norm_kv.bias[256]
 *       norm2.weight[256] norm2.bias[256]
 *   for j in 0..1,  prefix "transformer.decoder.layers.<j>.":
 *       for a in (self_attn, multihead_attn):
 *           a.q_proj.weight[256,256] a.q_proj.bias[256] a.k_proj.weight a.k_proj.bias a.v_proj.weight a.v_proj.bias
 *           a.merge.weight[256,256]
 *       mlp.0.weight[512,256] mlp.2.weight[256,512]
 *       norm1.weight norm1.bias norm2.weight norm2.bias norm3.weight norm3.bias        (each [256])
 *   query_embed1.weight[256] query_embed2.weight[256]
 *   tlbr_reg.0.weight[256,256] tlbr_reg.2.weight[4,256] tlbr_reg.2.bias[4]
 *   heatmap_conv.0.weight[256,256,3,3] heatmap_conv.0.bias[256] heatmap_conv.1.weight[256] heatmap_conv.1.bias[256]
 *   heatmap_conv.3.weight[256] heatmap_conv.3.bias[1]
 * (The decoder layers' unused q_proj/k_proj/v_proj/merge are not part of the blob.)
 *
 * max_h/max_w: PositionEncodingSine max_shape (cfg.NECK.MAX_SHAPE, src/config/default.py:25-28); 100,100.
 */
OETR_API int oetr_create(const float* weights, size_t n_floats, int weights_on_device,
                int attention_mode, int operand_precision, int max_h, int max_w,
                oetr_handle** out);

OETR_API int oetr_destroy(oetr_handle* h);

/* Bytes of caller-provided device scratch oetr_forward needs for this problem size. */
OETR_API int oetr_workspace_bytes(const oetr_handle* h, int batch, int hf1, int wf1, int hf2, int wf2, size_t* out);

/*
 * The hot path.  Replaces OETR.feature_correlation + center_estimation + size_regression +
 * box_tlbr_to_xyxy (src/model.py:240-250; forward's unclamped variant :193-211 when clamp == 0).
 *
 *   feat1 [batch,256,hf1,wf1], feat2 [batch,256,hf2,wf2]  fp32 NCHW device (output of input_proj2)
 *   img_h*, img_w*: model-input image sizes in pixels (src/model.py:230-233); stride = img_h / hf (integer)
 *   boxes1, boxes2 [batch,4] fp32 device, xyxy pixels
 *   dbg_* : nullable device outputs of the stage boundaries, for parity tests:
 *       dbg_hs     [2][batch][256]          decoder outputs hs1 | hs2
 *       dbg_memory [batch*L1 + batch*L2][256]  encoder outputs memory1 | memory2 (token-major)
 *       dbg_cxy    [2][batch][2]            soft-argmax centres (x,y)
 *       dbg_tlbr   [2][batch][4]            sigmoid(top,left,bottom,right)
 *   workspace: >= oetr_workspace_bytes(...) bytes, 256-byte aligned, device
 *   stream: a cudaStream_t (as void*); 0 = legacy default stream
 */
OETR_API int oetr_forward(oetr_handle* h,
                 const float* feat1, const float* feat2,
                 int batch, int hf1, int wf1, int hf2, int wf2,
                 int img_h1, int img_w1, int img_h2, int img_w2,
                 int clamp,
                 float* boxes1, float* boxes2,
                 float* dbg_hs, float* dbg_memory, float* dbg_cxy, float* dbg_tlbr,
                 void* workspace, size_t workspace_bytes, void* stream);

/* oetr_forward with the optional float masks of the reference's direct callers (forward_dummy(image1, image2, mask1,
 * mask2), src/model.py:229-250): mask1 [batch,hf1,wf1], mask2 [batch,hf2,wf2] fp32 device, or both NULL.  A position's
 * mask scales its phi(q), phi(k) and v in every encoder layer and in the decoder's cross-attention
 * (src/models/linear_attention.py:36-41, transformer.py:341-381), and heat-map logits where the mask is 0 are filled
 * with -1e9 before the softmax (src/model.py:167-171).  Linear attention only.  No shipped path of the reference
 * passes masks. */
OETR_API int oetr_forward_masked(oetr_handle* h,
                 const float* feat1, const float* feat2, const float* mask1, const float* mask2,
                 int batch, int hf1, int wf1, int hf2, int wf2,
                 int img_h1, int img_w1, int img_h2, int img_w2,
                 int clamp,
                 float* boxes1, float* boxes2,
                 float* dbg_hs, float* dbg_memory, float* dbg_cxy, float* dbg_tlbr,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Sub-batch scheduling of the FP16 path (default: automatic, about 56 encoder tiles per sub-batch = 8 pairs at
 * 640x640; env OETR_CHUNK_PAIRS overrides at create; < 0 = automatic): a batch larger
 * than `pairs_per_chunk` is cut into up to 8 balanced sub-batches that run on handle-owned streams, forked from and
 * joined to the caller's stream by events (no host synchronisation), so that free SMs are back-filled across
 * sub-batches and, in oetr_forward_host, copies overlap compute.  Results do not depend on it (every pair is
 * independent).  0 disables.  Call before oetr_workspace_bytes: the workspace requirement depends on it. */
OETR_API int oetr_set_chunk_pairs(oetr_handle* h, int pairs_per_chunk);

/* Number of kernels the last oetr_forward on this handle launched (bench.py's gpu_launches). */
OETR_API int oetr_last_launch_count(const oetr_handle* h);

/* Measurement hook (bench.py's roofline leg): when enabled, every launch of the dominant kernel of the FP16 path
 * (k_enc with a query phase, one per encoder layer; the profiler pass runs the batch unsplit on one stream) is bracketed by CUDA events on the launching stream.  oetr_profile_read
 * synchronises, returns the average launch duration since the last read and the number of launches, and resets. */
OETR_API int oetr_profile_enable(oetr_handle* h, int enable);
OETR_API int oetr_profile_read(oetr_handle* h, float* avg_ms, int* n_launches);

/* Synchronises the device and reports (and clears) asynchronous failures of earlier oetr_forward calls on this
 * handle: the tcgen05 kernels bound every mbarrier wait, so a pipeline protocol error surfaces here as
 * OETR_E_CUDA instead of hanging the GPU.  Intended for tests and debugging; not needed on the hot path. */
OETR_API int oetr_poll_error(oetr_handle* h);

/* Host-buffer entry points, for hosts without a CUDA runtime binding of their own (the e2e path of bench.py).
 *
 * oetr_forward_host_submit queues  host feats (pinned or pageable) -> device, forward, boxes -> pinned landing
 * buffer  on the handle's own streams, ordered after the work already queued on `stream`, and returns a ticket
 * without waiting.  Up to 4 requests may be in flight (a fifth submit fails with OETR_E_ARG until one is waited
 * for), so that the copies of request i+1 overlap the compute of request i.  The feature buffers must stay valid
 * and unchanged until the matching wait returns.  Device staging is owned by the handle and grown on demand
 * (these are the only entry points that allocate).
 * oetr_forward_host_wait blocks until that request has finished, copies its boxes to boxes1_host / boxes2_host
 * [batch,4] and reports a device-side failure of the request, if any.
 * oetr_forward_host = submit + wait. */
OETR_API int oetr_forward_host_submit(oetr_handle* h,
                      const float* feat1_host, const float* feat2_host,
                      int batch, int hf1, int wf1, int hf2, int wf2,
                      int img_h1, int img_w1, int img_h2, int img_w2,
                      int clamp, void* stream, int* ticket);
OETR_API int oetr_forward_host_wait(oetr_handle* h, int ticket, float* boxes1_host, float* boxes2_host);
OETR_API int oetr_forward_host(oetr_handle* h,
                      const float* feat1_host, const float* feat2_host,
                      int batch, int hf1, int wf1, int hf2, int wf2,
                      int img_h1, int img_w1, int img_h2, int img_w2,
                      int clamp, float* boxes1_host, float* boxes2_host, void* stream);

/* tcgen05 / TMEM / bulk-TMA self-test of the building blocks the FP16 path relies on (descriptor encodings,
 * swizzled operand images, TMEM accumulate).  Writes one max-abs-error per sub-test into errs[0..n_errs).
 * Returns OETR_OK when the kernels ran (inspect errs for the numeric outcome). */
OETR_API int oetr_selftest_tcgen05(float* errs_host, int n_errs);

/* The overlap head alone, for callers that use the reference's stage-wise methods (OETR.center_estimation,
 * src/model.py:145-186, OETR.size_regression :188-191, box_tlbr_to_xyxy src/models/utils.py:16-28) with their own
 * hs / memory tensors: memory1 [batch*hf1*wf1][256], memory2, hs1 [batch][256], hs2 (device, fp32, token-major),
 * optional masks [batch][hf*wf] -> boxes1/boxes2 [batch][4], cxy [2][batch][2] (nullable), tlbr [2][batch][4]
 * (nullable).  fp32 CUDA-core kernels for handles of either precision; workspace as for oetr_forward. */
OETR_API int oetr_head_forward(oetr_handle* h, const float* memory1, const float* memory2, const float* hs1, const float* hs2,
                               const float* mask1, const float* mask2, int batch, int hf1, int wf1, int hf2, int wf2,
                               int img_h1, int img_w1, int img_h2, int img_w2, int clamp, float* boxes1, float* boxes2,
                               float* cxy, float* tlbr, void* workspace, size_t workspace_bytes, void* stream);

/* ---- boxes of every rank on every rank (BASELINE configs[2]) ------------------------------------------
 * One process per GPU.  Instead of a collective, every rank stores its [pairs][2][4] boxes straight into its peers'
 * communication buffers over NVLink (CUDA IPC mapped peer memory, one tiny stream-ordered kernel) and publishes a
 * per-(slot, source rank) step flag; oetr_gather_collect waits (on the device, stream-ordered) for the flags of all
 * ranks of the oldest outstanding step and copies the gathered [world*pairs][2][4] boxes (global pair order: rank
 * major) to `all_boxes`.  Up to `slots` steps may be in flight; ranks never rendezvous.  Every submit must be
 * followed by a collect on every rank (a peer that stops collecting makes its writers trap after ~10 s).
 *   create : allocates this rank's buffer, returns its CUDA IPC handle (OETR_IPC_HANDLE_BYTES bytes)
 *   connect: all_handles = the handles of ranks 0..world-1, concatenated (exchange them with any host transport,
 *            e.g. torch.distributed.all_gather_object); world == 1 needs no connect
 * Replaces the reference-side nothing: the reference has no inference-time communication (train.py:59-74 is DDP). */
#define OETR_IPC_HANDLE_BYTES 64
typedef struct oetr_gather oetr_gather;
OETR_API int oetr_gather_create(int world, int rank, int pairs_per_rank, int slots, oetr_gather** out, void* ipc_handle_out);
OETR_API int oetr_gather_connect(oetr_gather* g, const void* all_handles);
OETR_API int oetr_gather_submit(oetr_gather* g, const float* boxes1, const float* boxes2, void* stream);
OETR_API int oetr_gather_collect(oetr_gather* g, float* all_boxes, void* stream);
OETR_API int oetr_gather_destroy(oetr_gather* g);
OETR_API const char* oetr_gather_last_error(void);

/* ---- the neck: what the reference runs between the backbone and the hot path ----------------------
 * Replaces, in OETR.feature_extraction (src/model.py:116-124): input_proj (1x1 conv 1024 -> 256, src/model.py:45-47),
 * PatchMerging (LayerNorm over channels + three stride-2 convolutions k = 4/8/16, padding (k-2)/2, 256 -> 256/128/128,
 * concatenated; src/models/backbone.py:28-67) and input_proj2 (1x1 conv 512 -> 256, src/model.py:48-50), for ONE image
 * set: backbone_out [n][1024][height][width] fp32 NCHW (ResNet-50 layer3) -> feat_out [n][256][height/2][width/2] fp32
 * NCHW (what oetr_forward reads).  tcgen05 kernels, single fp16 operands with fp32 accumulation; the convolution operands
 * are fetched with TMA tensor loads.  Stream-ordered, no allocation; workspace from oetr_neck_workspace_bytes.
 * Packed weights (fp32, host pointer), each tensor as torch's state_dict stores it:
 *   input_proj.weight[256,1024,1,1] input_proj.bias[256] patchmerging.norm.weight[256] patchmerging.norm.bias[256]
 *   patchmerging.reductions.0.weight[256,256,4,4] .0.bias[256] .1.weight[128,256,8,8] .1.bias[128]
 *   .2.weight[128,256,16,16] .2.bias[128] input_proj2.weight[256,512,1,1] input_proj2.bias[256]      (11 929 088 floats)
 * height, width in 2..200 (the position table of the hot path allows feature maps up to 100 x 100). */
typedef struct oetr_neck oetr_neck;
OETR_API size_t oetr_neck_packed_weight_count(void);
OETR_API int oetr_neck_create(const float* weights_host, size_t n_floats, oetr_neck** out);
OETR_API int oetr_neck_destroy(oetr_neck* h);
OETR_API int oetr_neck_workspace_bytes(const oetr_neck* h, int n_images, int height, int width, size_t* out);
OETR_API int oetr_neck_forward(oetr_neck* h, const float* backbone_out, int n_images, int height, int width, float* feat_out,
                               void* workspace, size_t workspace_bytes, void* stream);
OETR_API int oetr_neck_last_launch_count(const oetr_neck* h);
/* host-only: the convolution tiling of a problem on a GPU of `sms` SMs: out5 = {tiles, rows per tile (<= 128), output rows
 * per tile, images per tile, 100 * split-K parts of the k16 items (tile pairs) + split-K parts of the k8 + k4 items} */
OETR_API int oetr_neck_geometry(int n_images, int height, int width, int sms, int* out5);
OETR_API const char* oetr_neck_last_error(void);

/* ---- post-box plumbing on the device -------------------------------------------------------------
 * Replaces the host round trip of dloc/core/utils/utils.py:510-564 `tensor_overlap_crop` (called from evaluation.py:104-111):
 * image[0, :, y0:y1, x0:x1] -> * 255 -> cv2.resize(float32, (new_w, new_h), INTER_CUBIC) -> / 255.  One job = one resize
 * of one crop; all jobs of a call run in ONE kernel launch (up to 32 per launch), stream-ordered, no allocation.
 * src = the whole image [channels][src_h][src_w] fp32 (device); the crop is [y0, min(y1, src_h)) x [x0, min(x1, src_w))
 * like Python slicing; dst [channels][new_h][new_w] fp32 (device).  flags: OETR_CROP_MUL255 scales the source by 255
 * before interpolating, OETR_CROP_DIV255 divides the result by 255 (a two-pass resize sets MUL on the first pass and DIV on
 * the second).  Bicubic arithmetic = cv2 4.13 INTER_CUBIC on float32 (A = -0.75, clamped taps); results agree with cv2
 * to float32 rounding (2e-4 on the 0..255 scale). */
#define OETR_CROP_MUL255 1
#define OETR_CROP_DIV255 2
typedef struct oetr_crop_job {
    const float* src;
    float* dst;
    int channels, src_h, src_w;
    int x0, y0, x1, y1;
    int new_w, new_h;
    int flags;
} oetr_crop_job;
OETR_API int oetr_crop_resize(const oetr_crop_job* jobs, int n_jobs, void* stream);
OETR_API const char* oetr_crop_last_error(void);

/* ---- SuperGlue's two hot operators ---------------------------------------------------------------
 * oetr_sg_attention replaces `attention(query, key, value)` of third_party/SuperGluePretrainedNetwork/models/superglue.py:86-90
 * as called by MultiHeadedAttention.forward (:100-108): query [batch][256][n], key / value [batch][256][m] fp32 device
 * tensors in the reference's Conv1d layout, channel c = d * 4 + head (4 heads x 64 dims) -> out [batch][256][n]
 * = softmax_m(Q_h^T K_h / 8) V_h per head, online softmax (no [n, m] matrix in memory); `mode` selects the arithmetic.
 * oetr_sg_optimal_transport replaces `log_optimal_transport(scores, alpha, iters)` (:150-184): scores [batch][m][n] fp32,
 * alpha = bin_score -> out [batch][m+1][n+1] (log assignment matrix incl. dustbins, multiplied by m + n like the
 * reference).  workspace: oetr_sg_transport_workspace_bytes.  Both are stream-ordered and allocate nothing. */
#define OETR_SG_TENSOR 0   /* QK^T and PV on tcgen05, 3-term split fp16 operands, fp32 accumulation in TMEM (default) */
#define OETR_SG_FP32   1   /* the same operator in fp32 on the CUDA cores */
OETR_API int oetr_sg_attention(const float* query, const float* key, const float* value, float* out, int batch, int n, int m,
                               int mode, void* stream);
OETR_API size_t oetr_sg_transport_workspace_bytes(int batch, int m, int n);
OETR_API int oetr_sg_optimal_transport(const float* scores, float alpha, int iters, float* out, int batch, int m, int n,
                                       void* workspace, size_t workspace_bytes, void* stream);
OETR_API const char* oetr_sg_last_error(void);

/* Measurement aid: device-side accumulators of the tcgen05 kernels (per-tile MMA-lane busy / wait cycles, wall
 * nanoseconds per tile, row-warp stage durations; one atomicAdd per tile, no host synchronisation).  Switched on by
 * OETR_TIMING=1 in the environment or by oetr_debug_cycles(NULL, -1, 1) (off: (NULL, -1, 0)).  With n > 0: copies up
 * to n accumulators to out, optionally resets them, returns the number copied (0 when off); synchronises the device.
 * bench.py derives the dominant kernel's per-SM tile time inside the timed region from them. */
OETR_API int oetr_debug_cycles(unsigned long long* out, int n, int reset);

/* Host-only (no GPU needed) consistency check of the encoder's tile geometry for a problem size: every token is one
 * row of exactly one 128-token tile and the partial attention summaries gathered per image are exactly those of the
 * tiles holding its rows.  *flat_tiles = number of tiles of the flat tiling, 0 when per-image tiles are used (maps
 * under 128 tokens). */
OETR_API int oetr_selftest_geometry(int batch, int hf1, int wf1, int hf2, int wf2, int* flat_tiles);

OETR_API const char* oetr_last_error(void);
OETR_API int oetr_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif  /* OETR_B200_H_ */
