// C ABI of liboetr_b200.so (declared in include/oetr_b200.h): handle management, workspace carving and the
// stream-ordered orchestration of the hot path.  No torch types, no exceptions across the boundary.
#include "../../include/oetr_b200.h"
#include "oetr_common.cuh"
#include "tc_path.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

using namespace oetr;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(e_ == cudaErrorMemoryAllocation ? OETR_E_NOMEM : OETR_E_CUDA, "%s: %s (%s:%d)", \
                        #call, cudaGetErrorString(e_), __FILE__, __LINE__);                          \
    } while (0)

// Sub-batch scheduling of the FP16 path.  Pairs are independent, and one k_enc CTA owns a whole SM, so
// a batch of 256 tiles runs as two waves on 148 SMs with the second wave 27 % empty and every small kernel (k_fold,
// k_decoder, ...) leaving most of the GPU idle.  The batch is therefore cut into sub-batches of `chunk_pairs` pairs,
// each with its own workspace slice and its own stream: the block scheduler back-fills free SMs with CTAs of
// whichever sub-batch is ready (different layers interleave), and in oetr_forward_host the H2D copy of sub-batch
// c+1 overlaps the compute of sub-batch c.  Fork/join is by events on the caller's stream: no host synchronisation.
constexpr int MAX_CHUNKS = 8;
constexpr int HOST_SLOTS = 4;

// CUDA graph of the ~25 kernel launches of one sub-batch of a host request.  The host path owns every device buffer
// a sub-batch touches (staging, workspace, boxes), so the launch sequence of a (slot, sub-batch) is identical from
// request to request as long as the problem is: it is captured on its second use and replayed with one
// cudaGraphLaunch afterwards (the host thread then issues ~8 driver calls per sub-batch instead of ~35, which keeps
// the end-to-end rate from being bound by launch overhead).
struct ChunkGraph {
    cudaGraphExec_t exec = nullptr;
    int key[11] = {};               // Bc, geometry, image sizes, clamp
    const void* ptr[7] = {};        // feat1, feat2, boxes1, boxes2, workspace, position rows
    int uses = 0, launches = 0;
};

// one in-flight request of oetr_forward_host_submit: device staging, pinned landing buffers, completion events
struct HostSlot {
    ChunkGraph cg[MAX_CHUNKS];
    float *feat1 = nullptr, *feat2 = nullptr, *boxes = nullptr;
    float* boxes_pin = nullptr;    // pinned host landing buffer of the boxes (a D2H copy into pageable memory would
                                   // block the launching thread and serialise the sub-batches)
    int* flag_pin = nullptr;       // [MAX_CHUNKS] device timeout flag as seen at the end of each sub-batch
    void* ws = nullptr;
    size_t feat1_n = 0, feat2_n = 0, boxes_n = 0, boxes_pin_n = 0, ws_n = 0;
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_CHUNKS] = {};
    cudaStream_t aux[MAX_CHUNKS] = {};     // the slot's own sub-batch streams: requests in flight run side by side
    int n_join = 0, batch = 0, ticket = -1;
    bool busy = false;
};

struct oetr_handle {
    int attn_mode = 0, prec = 0, max_h = 0, max_w = 0, device = 0;
    int chunk_pairs = -1;      // pairs per sub-batch (0: never split, < 0: chosen from the geometry); oetr_set_chunk_pairs
    cudaStream_t aux[MAX_CHUNKS] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_CHUNKS] = {};
    std::mutex mu;             // fork/launch/join section (the events are shared by all callers of this handle)
    WLayout L;
    float* d_w = nullptr;      // packed fp32 weights (canonical order)
    std::vector<float> h_w;    // host copy (LayerNorm vectors / biases travel to the kernels as launch parameters)
    float* d_w9 = nullptr;     // heatmap_conv.0.weight repacked per tap: [9][256 out][256 in]
    float* d_pe = nullptr;     // PositionEncodingSine table, channel-last: [max_h][max_w][256]
    TcWeights tc;              // fp16 UMMA operand images (FP16 path only)
    KernelProfiler prof;       // optional CUDA-event brackets around the dominant kernel
    int* d_flag = nullptr;     // raised by a device-side mbarrier wait that timed out (protocol bug guard)
    // FP16 path: tile-blocked positional rows per feature-map geometry.  Entries are IMMUTABLE once generated (a forward
    // on another stream may still be reading them): a new geometry gets a new buffer, and every forward makes its
    // stream wait for the generating kernel through the entry's event.
    struct PosEntry { int hf = 0, wf = 0; float* rows = nullptr; cudaEvent_t ready = nullptr; };
    std::vector<PosEntry> pos_cache;
    int last_launches = 0;
    // staging owned by the handle for the host-buffer entry points only: HOST_SLOTS requests can be in flight
    HostSlot slot[HOST_SLOTS];
    unsigned next_ticket = 0;
};

// ---------------------------------------------------------------------------------------------------------
// small setup kernels
// ---------------------------------------------------------------------------------------------------------
__global__ void k_repack_conv(const float* __restrict__ w, float* __restrict__ w9) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // over [tap][o][c]
    if (idx >= 9 * C * C) return;
    const int c = idx % C, o = (idx / C) % C, tap = idx / (C * C);
    w9[idx] = w[((size_t)o * C + c) * 9 + tap];
}
// token-major positional rows of an (hf,wf) map: out[l][:] = pe[l / wf][l % wf][:]   (models/utils.py:200-205)
__global__ void k_gather_pos(const float* __restrict__ pe, int max_w, int wf, int L, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L * (C / 4)) return;
    const int c4 = idx % (C / 4), l = idx / (C / 4);
    reinterpret_cast<float4*>(out)[idx] =
        reinterpret_cast<const float4*>(pe)[((size_t)(l / wf) * max_w + (l % wf)) * (C / 4) + c4];
}

// PositionEncodingSine (src/models/utils.py:185-198) incl. the precedence quirk: div_term = exp(-2k), k=0..63.
static void build_pe_host(std::vector<float>& pe, int max_h, int max_w) {
    pe.assign((size_t)max_h * max_w * C, 0.f);
    const float factor = std::floor((float)(-std::log(10000.0) / C) / 2.0f);    // == -1.0f
    for (int k = 0; k < C / 4; ++k) {
        const float div = (float)std::exp((double)((float)(2 * k) * factor));   // fp32 exp of fp32 arg
        for (int y = 0; y < max_h; ++y)
            for (int x = 0; x < max_w; ++x) {
                float* p = &pe[((size_t)y * max_w + x) * C + 4 * k];
                const float ax = (float)(x + 1) * div, ay = (float)(y + 1) * div;   // fp32 products
                p[0] = (float)std::sin((double)ax);
                p[1] = (float)std::cos((double)ax);
                p[2] = (float)std::sin((double)ay);
                p[3] = (float)std::cos((double)ay);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------
// workspace carving
// ---------------------------------------------------------------------------------------------------------
namespace {
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* b) : base(static_cast<char*>(b)) {}
    template <typename T> T* take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};
struct Workspace {
    float *X, *T, *Q, *KF, *V, *O, *H, *kvs, *pos;
    float *dt, *du, *dqk, *dq, *dk, *dv, *dob, *dh, *dkvs;
    TcWorkspace tc;
    size_t bytes;
};
Workspace carve(void* base, const oetr_handle* h, int B, int L1, int L2) {
    Workspace w{};
    Carver c(base);
    const size_t R = (size_t)B * (L1 + L2);
    w.X = c.take<float>(R * C); w.T = c.take<float>(R * C); w.Q = c.take<float>(R * C);
    w.KF = c.take<float>(R * C); w.V = c.take<float>(R * C); w.O = c.take<float>(R * C);
    w.H = c.take<float>(R * FF);
    w.kvs = c.take<float>((size_t)2 * B * KVS);
    w.pos = c.take<float>((size_t)(L1 + L2) * C);
    const size_t D = (size_t)2 * B * C;
    w.dt = c.take<float>(D); w.du = c.take<float>(D); w.dqk = c.take<float>(D); w.dq = c.take<float>(D);
    w.dk = c.take<float>(D); w.dv = c.take<float>(D); w.dob = c.take<float>(D); w.dh = c.take<float>(2 * D);
    w.dkvs = c.take<float>((size_t)2 * B * KVS);
    if (h->prec == OETR_PREC_FP16) tc_carve(c.off, base, B, L1, L2, w.tc, h->attn_mode == OETR_ATTN_FULL);
    w.bytes = (c.off + 255) & ~size_t(255);
    return w;
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------
// fp32 orchestration
// ---------------------------------------------------------------------------------------------------------
static void encoder_fp32(const oetr_handle* h, const Workspace& w, int B, int L1, int L2, const float* mask1,
                         const float* mask2, cudaStream_t s, LaunchCounter& lc) {
    const float* W = h->d_w;
    const int R1 = B * L1, R2 = B * L2, R = R1 + R2;
    float* const pos1 = w.pos;
    float* const pos2 = w.pos + (size_t)L1 * C;
    auto ln_both = [&](const float* g, const float* b, bool with_pos, float* out) {
        ln_pos(w.X, g, b, with_pos ? pos1 : nullptr, L1, out, R1, s, lc);
        ln_pos(w.X + (size_t)R1 * C, g, b, with_pos ? pos2 : nullptr, L2, out + (size_t)R1 * C, R2, s, lc);
    };
    const bool linear = h->attn_mode == OETR_ATTN_LINEAR;
    // q / kv masks of LinearAttention: a row's own mask scales its phi(q), phi(k) and v (cross layers read the
    // partner's summaries, which were built from the partner's masked rows)
    auto mask_rows = [&](float* buf) {
        row_scale(buf, mask1, R1, s, lc);
        row_scale(buf + (size_t)R1 * C, mask2, R2, s, lc);
    };
    for (int i = 0; i < N_ENC; ++i) {
        const EncW& e = h->L.enc[i];
        // key/value side: ONE normalised source (+pos) feeds k_proj and v_proj (transformer.py:119-126)
        ln_both(W + e.lnkv_g, W + e.lnkv_b, true, w.T);
        gemm_nt(w.T, C, W + e.wk, C, nullptr, w.KF, C, R, C, C, linear ? ACT_ELU1 : ACT_NONE, 0, s, lc);
        gemm_nt(w.T, C, W + e.wv, C, nullptr, w.V, C, R, C, C, ACT_NONE, 0, s, lc);
        if (linear) {
            mask_rows(w.KF);
            mask_rows(w.V);
            kv_reduce(w.KF, w.V, w.kvs, B, L1, s, lc);
            kv_reduce(w.KF + (size_t)R1 * C, w.V + (size_t)R1 * C, w.kvs + (size_t)B * KVS, B, L2, s, lc);
        }
        // query side
        ln_both(W + e.lnq_g, W + e.lnq_b, true, w.T);
        gemm_nt(w.T, C, W + e.wq, C, nullptr, w.Q, C, R, C, C, linear ? ACT_ELU1 : ACT_NONE, 0, s, lc);
        const bool cross = (i & 1) != 0;   // ['self','cross']*4; cross layers read the partner's PRE-update state
        if (linear) {
            mask_rows(w.Q);
            const float* kv_for_1 = w.kvs + (cross ? (size_t)B * KVS : 0);
            const float* kv_for_2 = w.kvs + (cross ? 0 : (size_t)B * KVS);
            linattn_apply(w.Q, kv_for_1, w.O, R1, L1, s, lc);
            linattn_apply(w.Q + (size_t)R1 * C, kv_for_2, w.O + (size_t)R1 * C, R2, L2, s, lc);
        } else {
            const size_t o2 = (size_t)R1 * C;
            if (!cross) {
                full_attention(w.Q, w.KF, w.V, w.O, B, L1, L1, s, lc);
                full_attention(w.Q + o2, w.KF + o2, w.V + o2, w.O + o2, B, L2, L2, s, lc);
            } else {
                full_attention(w.Q, w.KF + o2, w.V + o2, w.O, B, L1, L2, s, lc);
                full_attention(w.Q + o2, w.KF, w.V, w.O + o2, B, L2, L1, s, lc);
            }
        }
        gemm_nt(w.O, C, W + e.wm, C, nullptr, w.X, C, R, C, C, ACT_NONE, 1, s, lc);          // x += merge(att)
        ln_both(W + e.ln2_g, W + e.ln2_b, false, w.T);
        gemm_nt(w.T, C, W + e.w1, C, nullptr, w.H, FF, R, FF, C, ACT_GELU, 0, s, lc);
        gemm_nt(w.H, FF, W + e.w2, FF, nullptr, w.X, C, R, C, FF, ACT_NONE, 1, s, lc);        // x += mlp
    }
}

// Query decoder (transformer.py:224-284,361-381) for both image sets stacked as rows [0,B) | [B,2B).
// `decoder_kv` supplies the cross-attention summaries of decoder layer j in w.dkvs.
template <class KvFn>
static void decoder_fp32(const oetr_handle* h, const Workspace& w, int B, cudaStream_t s, LaunchCounter& lc,
                         KvFn decoder_kv) {
    const float* W = h->d_w;
    const int D2 = 2 * B;
    const float* qe = W + h->L.qe1;                       // qe1 | qe2 are adjacent in the blob
    cudaMemsetAsync(w.dt, 0, (size_t)D2 * C * sizeof(float), s);
    for (int j = 0; j < N_DEC; ++j) {
        const DecW& d = h->L.dec[j];
        // self-attention over the single query token
        ln_pos(w.dt, W + d.ln1_g, W + d.ln1_b, nullptr, 1, w.du, D2, s, lc);
        ln_pos(w.du, nullptr, nullptr, qe, -B, w.dqk, D2, s, lc);               // u + query_embed{1,2}
        gemm_nt(w.dqk, C, W + d.sa.wq, C, W + d.sa.bq, w.dq, C, D2, C, C, ACT_ELU1, 0, s, lc);
        gemm_nt(w.dqk, C, W + d.sa.wk, C, W + d.sa.bk, w.dk, C, D2, C, C, ACT_ELU1, 0, s, lc);
        gemm_nt(w.du, C, W + d.sa.wv, C, W + d.sa.bv, w.dv, C, D2, C, C, ACT_NONE, 0, s, lc);
        kv_reduce(w.dk, w.dv, w.dkvs, D2, 1, s, lc);
        linattn_apply(w.dq, w.dkvs, w.dob, D2, 1, s, lc);
        gemm_nt(w.dob, C, W + d.sa.wm, C, nullptr, w.dt, C, D2, C, C, ACT_NONE, 1, s, lc);
        // cross-attention into the encoder memory: k = memory+pos, v = memory (no LN, no pos on v)
        ln_pos(w.dt, W + d.ln2_g, W + d.ln2_b, qe, -B, w.dqk, D2, s, lc);
        gemm_nt(w.dqk, C, W + d.ca.wq, C, W + d.ca.bq, w.dq, C, D2, C, C, ACT_ELU1, 0, s, lc);
        decoder_kv(j);
        linattn_apply(w.dq, w.dkvs, w.dob, D2, 1, s, lc);
        gemm_nt(w.dob, C, W + d.ca.wm, C, nullptr, w.dt, C, D2, C, C, ACT_NONE, 1, s, lc);
        // feed-forward
        ln_pos(w.dt, W + d.ln3_g, W + d.ln3_b, nullptr, 1, w.du, D2, s, lc);
        gemm_nt(w.du, C, W + d.w1, C, nullptr, w.dh, FF, D2, FF, C, ACT_RELU, 0, s, lc);
        gemm_nt(w.dh, FF, W + d.w2, FF, nullptr, w.dt, C, D2, C, FF, ACT_NONE, 1, s, lc);
    }
}

static void decoder_kv_fp32(const oetr_handle* h, const Workspace& w, int j, int B, int L1, int L2,
                            const float* mask1, const float* mask2, cudaStream_t s, LaunchCounter& lc) {
    const float* W = h->d_w;
    const DecW& d = h->L.dec[j];
    const int R1 = B * L1, R2 = B * L2, R = R1 + R2;
    ln_pos(w.X, nullptr, nullptr, w.pos, L1, w.T, R1, s, lc);
    ln_pos(w.X + (size_t)R1 * C, nullptr, nullptr, w.pos + (size_t)L1 * C, L2, w.T + (size_t)R1 * C, R2, s, lc);
    gemm_nt(w.T, C, W + d.ca.wk, C, W + d.ca.bk, w.KF, C, R, C, C, ACT_ELU1, 0, s, lc);
    gemm_nt(w.X, C, W + d.ca.wv, C, W + d.ca.bv, w.V, C, R, C, C, ACT_NONE, 0, s, lc);
    row_scale(w.KF, mask1, R1, s, lc); row_scale(w.KF + (size_t)R1 * C, mask2, R2, s, lc);      // memory_mask (transformer.py:361-381)
    row_scale(w.V, mask1, R1, s, lc);  row_scale(w.V + (size_t)R1 * C, mask2, R2, s, lc);
    kv_reduce(w.KF, w.V, w.dkvs, B, L1, s, lc);
    kv_reduce(w.KF + (size_t)R1 * C, w.V + (size_t)R1 * C, w.dkvs + (size_t)B * KVS, B, L2, s, lc);
}

// conv3x3 as 9 shifted GEMMs accumulating into Y (bias added by the first tap)
static void head_conv_fp32(const oetr_handle* h, const Workspace& w, int B, int hf1, int wf1, int hf2, int wf2,
                           const float* G, float* Gs, float* Y, cudaStream_t s, LaunchCounter& lc) {
    const int R1 = B * hf1 * wf1, R2 = B * hf2 * wf2, R = R1 + R2;
    for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        shift_tokens(G, Gs, B, hf1, wf1, dy, dx, s, lc);
        shift_tokens(G + (size_t)R1 * C, Gs + (size_t)R1 * C, B, hf2, wf2, dy, dx, s, lc);
        gemm_nt(Gs, C, h->d_w9 + (size_t)tap * C * C, C, tap == 0 ? h->d_w + h->L.hm_b0 : nullptr, Y, C, R, C, C,
                ACT_NONE, tap == 0 ? 0 : 1, s, lc);
    }
}

// ---------------------------------------------------------------------------------------------------------
// exported functions
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int oetr_abi_version(void) { return OETR_ABI_VERSION; }
const char* oetr_last_error(void) { return g_err; }
size_t oetr_packed_weight_count(void) { return make_layout().total; }

int oetr_create(const float* weights, size_t n_floats, int weights_on_device, int attention_mode,
                int operand_precision, int max_h, int max_w, oetr_handle** out) {
    if (!weights || !out) return fail(OETR_E_ARG, "oetr_create: null argument");
    *out = nullptr;
    if (attention_mode != OETR_ATTN_LINEAR && attention_mode != OETR_ATTN_FULL)
        return fail(OETR_E_ARG, "oetr_create: attention_mode %d not in {0 linear, 1 full}", attention_mode);
    if (operand_precision != OETR_PREC_FP32 && operand_precision != OETR_PREC_FP16)
        return fail(OETR_E_ARG, "oetr_create: operand_precision %d not in {0 fp32, 1 fp16}", operand_precision);
    if (max_h < 1 || max_w < 1 || max_h > 1024 || max_w > 1024)
        return fail(OETR_E_SHAPE, "oetr_create: max_shape (%d,%d) out of range", max_h, max_w);
    const WLayout L = make_layout();
    if (n_floats != L.total)
        return fail(OETR_E_ARG, "oetr_create: expected %zu packed floats, got %zu", L.total, n_floats);
    int dev = 0;
    CU(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(OETR_E_ARCH, "oetr_create: device %d is sm_%d%d; this library is sm_100a-only (no fallback)",
                    dev, prop.major, prop.minor);
    oetr_handle* h = new (std::nothrow) oetr_handle();
    if (!h) return fail(OETR_E_NOMEM, "oetr_create: host allocation failed");
    h->attn_mode = attention_mode; h->prec = operand_precision; h->max_h = max_h; h->max_w = max_w;
    h->device = dev; h->L = L;
    auto bail = [&](int code) { oetr_destroy(h); return code; };
#define CUH(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return bail(fail(e_ == cudaErrorMemoryAllocation ? OETR_E_NOMEM : OETR_E_CUDA, "%s: %s", #call, \
                             cudaGetErrorString(e_)));                                              \
    } while (0)
    CUH(cudaMalloc(&h->d_w, L.total * sizeof(float)));
    CUH(cudaMemcpy(h->d_w, weights, L.total * sizeof(float),
                   weights_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    h->h_w.resize(L.total);
    CUH(cudaMemcpy(h->h_w.data(), h->d_w, L.total * sizeof(float), cudaMemcpyDeviceToHost));
    CUH(cudaMalloc(&h->d_w9, (size_t)9 * C * C * sizeof(float)));
    k_repack_conv<<<(9 * C * C + 255) / 256, 256>>>(h->d_w + L.hm_w0, h->d_w9);
    std::vector<float> pe;
    build_pe_host(pe, max_h, max_w);
    CUH(cudaMalloc(&h->d_pe, pe.size() * sizeof(float)));
    CUH(cudaMemcpy(h->d_pe, pe.data(), pe.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUH(cudaMalloc(&h->d_flag, sizeof(int)));
    CUH(cudaMemset(h->d_flag, 0, sizeof(int)));
    if (const char* env = getenv("OETR_CHUNK_PAIRS")) h->chunk_pairs = atoi(env);
    CUH(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < MAX_CHUNKS; ++i) {
        CUH(cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking));
        CUH(cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
    }
    for (HostSlot& sl : h->slot) {
        CUH(cudaEventCreateWithFlags(&sl.ev_fork, cudaEventDisableTiming));
        for (int i = 0; i < MAX_CHUNKS; ++i) {
            CUH(cudaEventCreateWithFlags(&sl.ev_join[i], cudaEventDisableTiming));
            CUH(cudaStreamCreateWithFlags(&sl.aux[i], cudaStreamNonBlocking));
        }
        CUH(cudaMallocHost(&sl.flag_pin, MAX_CHUNKS * sizeof(int)));
    }
    if (operand_precision == OETR_PREC_FP16) {
        char msg[256] = "";
        if (tc_prepare_weights(h->d_w, h->d_w9, L, h->tc, msg, sizeof(msg)) != 0)
            return bail(fail(OETR_E_CUDA, "oetr_create: %s", msg));
    }
    CUH(cudaDeviceSynchronize());
#undef CUH
    *out = h;
    return OETR_OK;
}

int oetr_destroy(oetr_handle* h) {
    if (!h) return OETR_OK;
    cudaFree(h->d_w); cudaFree(h->d_w9); cudaFree(h->d_pe); cudaFree(h->d_flag);
    for (auto& e : h->pos_cache) { cudaFree(e.rows); if (e.ready) cudaEventDestroy(e.ready); }
    tc_free_weights(h->tc);
    for (cudaEvent_t e : h->prof.ev) cudaEventDestroy(e);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (int i = 0; i < MAX_CHUNKS; ++i) {
        if (h->aux[i]) cudaStreamDestroy(h->aux[i]);
        if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    }
    for (HostSlot& sl : h->slot) {
        if (sl.busy) for (int c = 0; c < sl.n_join; ++c) cudaEventSynchronize(sl.ev_join[c]);
        for (ChunkGraph& g : sl.cg) if (g.exec) cudaGraphExecDestroy(g.exec);
        cudaFree(sl.feat1); cudaFree(sl.feat2); cudaFree(sl.boxes); cudaFree(sl.ws);
        cudaFreeHost(sl.boxes_pin); cudaFreeHost(sl.flag_pin);
        if (sl.ev_fork) cudaEventDestroy(sl.ev_fork);
        for (int i = 0; i < MAX_CHUNKS; ++i) {
            if (sl.ev_join[i]) cudaEventDestroy(sl.ev_join[i]);
            if (sl.aux[i]) cudaStreamDestroy(sl.aux[i]);
        }
    }
    delete h;
    return OETR_OK;
}

static int check_shapes(const oetr_handle* h, int batch, int hf1, int wf1, int hf2, int wf2) {
    if (!h) return fail(OETR_E_ARG, "null handle");
    if (batch < 0) return fail(OETR_E_ARG, "batch %d < 0", batch);
    if (hf1 < 1 || wf1 < 1 || hf2 < 1 || wf2 < 1 || hf1 > h->max_h || hf2 > h->max_h || wf1 > h->max_w ||
        wf2 > h->max_w)
        return fail(OETR_E_SHAPE, "feature maps (%d,%d),(%d,%d) outside [1,(%d,%d)] (PositionEncodingSine max_shape)",
                    hf1, wf1, hf2, wf2, h->max_h, h->max_w);
    return OETR_OK;
}

// sub-batch sizes of a batch (balanced; 1 chunk = no split)
static int chunk_plan(const oetr_handle* h, int B, int L1, int L2, int* sizes) {
    int n = 1;
    int cp = h->chunk_pairs;
    if (cp < 0) {
        // automatic: about 56 encoder tiles (128 tokens) per sub-batch -- three sub-batches fill the 148 SMs and a
        // fourth back-fills; 8 pairs at 640x640 (400 tokens per image), 5 at 840x840 (676)
        const int tiles16 = ((L1 + 15) / 16 + (L2 + 15) / 16 + 7) / 8 * 2;      // 2 x tiles per pair, flat tiling
        cp = (2 * 56 + tiles16 / 2) / (tiles16 > 0 ? tiles16 : 1);
        if (cp < 1) cp = 1;
    }
    if (h->prec == OETR_PREC_FP16 && cp > 0 && B > cp)
        n = (B + cp - 1) / cp;
    if (n > MAX_CHUNKS) n = MAX_CHUNKS;
    for (int i = 0; i < n; ++i) sizes[i] = B / n + (i < B % n ? 1 : 0);
    return n;
}
static size_t chunked_bytes(const oetr_handle* h, int B, int L1, int L2) {
    int sizes[MAX_CHUNKS];
    const int n = chunk_plan(h, B, L1, L2, sizes);
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += (carve(nullptr, h, sizes[i], L1, L2).bytes + 1023) & ~size_t(1023);
    return total;
}

int oetr_workspace_bytes(const oetr_handle* h, int batch, int hf1, int wf1, int hf2, int wf2, size_t* out) {
    if (!out) return fail(OETR_E_ARG, "oetr_workspace_bytes: null out");
    int rc = check_shapes(h, batch, hf1, wf1, hf2, wf2);
    if (rc) return rc;
    const size_t whole = carve(nullptr, h, batch, hf1 * wf1, hf2 * wf2).bytes;
    const size_t split = chunked_bytes(h, batch, hf1 * wf1, hf2 * wf2);
    *out = (whole > split ? whole : split) + 256;
    return OETR_OK;
}

int oetr_set_chunk_pairs(oetr_handle* h, int pairs_per_chunk) {
    if (!h) return fail(OETR_E_ARG, "null handle");
    h->chunk_pairs = pairs_per_chunk;
    return OETR_OK;
}

int oetr_last_launch_count(const oetr_handle* h) { return h ? h->last_launches : 0; }

int oetr_profile_enable(oetr_handle* h, int enable) {
    if (!h) return fail(OETR_E_ARG, "null handle");
    h->prof.on = enable != 0;
    h->prof.used = 0;
    return OETR_OK;
}

int oetr_profile_read(oetr_handle* h, float* avg_ms, int* n_launches) {
    if (!h || !avg_ms || !n_launches) return fail(OETR_E_ARG, "oetr_profile_read: null argument");
    CU(cudaDeviceSynchronize());
    double total = 0.0;
    int n = 0;
    for (size_t i = 0; i + 1 < h->prof.used; i += 2) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->prof.ev[i], h->prof.ev[i + 1]));
        total += ms;
        ++n;
    }
    *avg_ms = n ? (float)(total / n) : 0.f;
    *n_launches = n;
    h->prof.used = 0;
    return OETR_OK;
}

int oetr_poll_error(oetr_handle* h) {
    if (!h) return fail(OETR_E_ARG, "null handle");
    CU(cudaDeviceSynchronize());
    int flag = 0;
    CU(cudaMemcpy(&flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
        cudaMemset(h->d_flag, 0, sizeof(int));
        return fail(OETR_E_CUDA, "a device-side mbarrier wait timed out (pipeline protocol error); results are invalid");
    }
    return OETR_OK;
}

namespace {
// optional host endpoints of a forward (oetr_forward_host): features are copied H2D and boxes D2H on the stream
// that runs the (sub-)batch, so that with sub-batch scheduling the copies overlap the other sub-batches' compute
struct HostIO { const float *feat1, *feat2; float *boxes1, *boxes2; int* flags; ChunkGraph* graphs; };
// fork/join events of one forward.  join_caller: the caller's stream waits for the sub-batches (stream-ordered
// entry points); otherwise completion is observed through ev_join only (host submit/wait) and even an unsplit
// batch runs on a handle-owned stream, so that two requests forked from the same stream overlap
struct EventSet { cudaEvent_t fork; cudaEvent_t* join; bool join_caller; int* n_join; cudaStream_t* streams; };

struct FwdArgs {
    int hf1, wf1, hf2, wf2, img_h1, img_w1, img_h2, img_w2, clamp;
    float *dbg_hs, *dbg_memory, *dbg_cxy, *dbg_tlbr;
    const float *mask1, *mask2;     // nullable [batch][hf*wf] device masks of the WHOLE batch (sub-batches offset them)
    const float *post1 = nullptr, *post2 = nullptr;   // FP16 path: tile-blocked positional rows of the two geometries
};

// one (sub-)batch of B pairs on stream s with workspace slice w
int run_batch(oetr_handle* h, const Workspace& w, const float* feat1, const float* feat2, int B, const FwdArgs& a,
              float* boxes1, float* boxes2, bool profile, cudaStream_t s, LaunchCounter& lc) {
    const int hf1 = a.hf1, wf1 = a.wf1, hf2 = a.hf2, wf2 = a.wf2;
    const int L1 = hf1 * wf1, L2 = hf2 * wf2, R1 = B * L1, R2 = B * L2;
    const float* W = h->d_w;
    if (h->prec == OETR_PREC_FP32) {
        // positional rows for both geometries (PositionEncodingSine.forward slice, models/utils.py:200-205)
        k_gather_pos<<<(L1 * (C / 4) + 255) / 256, 256, 0, s>>>(h->d_pe, h->max_w, wf1, L1, w.pos); lc.n++;
        k_gather_pos<<<(L2 * (C / 4) + 255) / 256, 256, 0, s>>>(h->d_pe, h->max_w, wf2, L2, w.pos + (size_t)L1 * C); lc.n++;
        nchw_to_tokens(feat1, w.X, B, L1, s, lc);
        nchw_to_tokens(feat2, w.X + (size_t)R1 * C, B, L2, s, lc);
        encoder_fp32(h, w, B, L1, L2, a.mask1, a.mask2, s, lc);
        decoder_fp32(h, w, B, s, lc, [&](int j) { decoder_kv_fp32(h, w, j, B, L1, L2, a.mask1, a.mask2, s, lc); });
    } else {
        // tcgen05 encoder + decoder K/V summaries (memory stays tile-blocked in w.tc.xt; token-major copy only for
        // the debug output), fused fp32 decoder, tcgen05 heat-map convolution
        char msg[256] = "";
        const int erc = h->attn_mode == OETR_ATTN_FULL
            ? tc_encoder_full(h->tc, h->d_w, h->h_w.data(), h->L, w.tc, feat1, feat2, B, hf1, wf1, hf2, wf2, a.post1, a.post2,
                              a.dbg_memory ? w.X : nullptr, h->d_flag, s, lc, msg, sizeof(msg))
            : tc_encoder(h->tc, h->d_w, h->h_w.data(), h->L, w.tc, feat1, feat2, B, hf1, wf1, hf2, wf2, a.post1, a.post2,
                         a.mask1, a.mask2,
                         a.dbg_memory ? w.X : nullptr, h->d_flag, profile ? &h->prof : nullptr, s, lc, msg, sizeof(msg));
        if (erc != 0) return fail(OETR_E_CUDA, "oetr_forward: %s", msg);
        const HeadGeom hg{B, hf1, wf1, hf2, wf2, a.img_h1, a.img_w1, a.img_h2, a.img_w2, a.clamp, a.mask1, a.mask2};
        if (tc_decoder_head(h->tc, h->d_w, h->L, w.tc, hg, w.dt, w.O, boxes1, boxes2, a.dbg_cxy, a.dbg_tlbr, h->d_flag, s, lc,
                            msg, sizeof(msg)) != 0)
            return fail(OETR_E_CUDA, "oetr_forward: %s", msg);
    }
    // memory = encoder output (w.X), hs = decoder output (w.dt)
    if (a.dbg_memory) cudaMemcpyAsync(a.dbg_memory, w.X, (size_t)(R1 + R2) * C * sizeof(float), cudaMemcpyDeviceToDevice, s);
    if (a.dbg_hs) cudaMemcpyAsync(a.dbg_hs, w.dt, (size_t)2 * B * C * sizeof(float), cudaMemcpyDeviceToDevice, s);

    if (h->prec == OETR_PREC_FP32) {
        // head: heat = memory * <memory, hs>; conv3x3 (+bias) -> Y ; then the row-wise tail per image
        float *G = w.T, *Gs = w.Q, *Y = w.O;
        heat_scale(w.X, w.dt, G, R1, L1, s, lc);
        heat_scale(w.X + (size_t)R1 * C, w.dt + (size_t)B * C, G + (size_t)R1 * C, R2, L2, s, lc);
        head_conv_fp32(h, w, B, hf1, wf1, hf2, wf2, G, Gs, Y, s, lc);
        HeadParams p{};
        p.gn_g = W + h->L.hm_gn_g; p.gn_b = W + h->L.hm_gn_b; p.w3 = W + h->L.hm_w3; p.b3 = W + h->L.hm_b3;
        p.tl_w0 = W + h->L.tl_w0; p.tl_w2 = W + h->L.tl_w2; p.tl_b2 = W + h->L.tl_b2;
        p.batch = B; p.clamp = a.clamp;
        p.Y = Y; p.hs = w.dt; p.hf = hf1; p.wf = wf1; p.img_h = a.img_h1; p.img_w = a.img_w1;
        p.boxes = boxes1; p.dbg_cxy = a.dbg_cxy; p.dbg_tlbr = a.dbg_tlbr; p.mask = a.mask1;
        head_finalize(p, s, lc);
        p.Y = Y + (size_t)R1 * C; p.hs = w.dt + (size_t)B * C; p.hf = hf2; p.wf = wf2; p.img_h = a.img_h2; p.img_w = a.img_w2;
        p.boxes = boxes2; p.dbg_cxy = a.dbg_cxy ? a.dbg_cxy + 2 * B : nullptr; p.dbg_tlbr = a.dbg_tlbr ? a.dbg_tlbr + 4 * B : nullptr;
        p.mask = a.mask2;
        head_finalize(p, s, lc);
    }
    return OETR_OK;
}

int forward_core(oetr_handle* h, const float* feat1, const float* feat2, int batch, const FwdArgs& a_in, float* boxes1,
                 float* boxes2, void* workspace, size_t workspace_bytes, cudaStream_t s, const HostIO* hio,
                 const EventSet& es) {
    FwdArgs a = a_in;
    const int B = batch, L1 = a.hf1 * a.wf1, L2 = a.hf2 * a.wf2;
    int cur_dev = -1;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev != h->device)
        return fail(OETR_E_ARG, "oetr_forward: the handle lives on device %d but the current device is %d", h->device, cur_dev);
    LaunchCounter lc;
    const float* post[2] = {nullptr, nullptr};
    if (h->prec == OETR_PREC_FP16) {
        // positional rows, tile-blocked, cached per geometry (the only state a forward adds to the handle)
        const int geo[2][2] = {{a.hf1, a.wf1}, {a.hf2, a.wf2}};
        for (int k = 0; k < 2; ++k) {
            oetr_handle::PosEntry* hit = nullptr;
            for (auto& e : h->pos_cache) if (e.hf == geo[k][0] && e.wf == geo[k][1]) hit = &e;
            if (!hit) {
                if (h->pos_cache.size() >= 32) {               // many geometries: start over once nothing can be reading
                    CU(cudaDeviceSynchronize());
                    for (auto& e : h->pos_cache) { cudaFree(e.rows); cudaEventDestroy(e.ready); }
                    h->pos_cache.clear();
                    for (HostSlot& sl : h->slot) for (ChunkGraph& g : sl.cg) g.uses = 0, g.ptr[5] = g.ptr[6] = nullptr;
                }
                oetr_handle::PosEntry e;
                e.hf = geo[k][0]; e.wf = geo[k][1];
                const int Lk = e.hf * e.wf;
                CU(cudaMalloc(&e.rows, tc_pos_tile_floats(Lk) * sizeof(float)));
                CU(cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming));
                tc_pos_tiles(h->d_pe, h->max_w, e.wf, Lk, e.rows, s, lc);
                CU(cudaEventRecord(e.ready, s));
                h->pos_cache.push_back(e);
                hit = &h->pos_cache.back();
            }
            CU(cudaStreamWaitEvent(s, hit->ready, 0));         // no-op once the generating kernel has finished
            post[k] = hit->rows;
        }
        a.post1 = post[0]; a.post2 = post[1];
    }
    int sizes[MAX_CHUNKS];
    int nc = chunk_plan(h, B, L1, L2, sizes);
    // the debug taps are laid out for the whole batch and the kernel profiler brackets launches on one stream
    if (a.dbg_hs || a.dbg_memory || a.dbg_cxy || a.dbg_tlbr || h->prof.on) nc = 1;
    if (nc == 1 && es.join_caller) {
        const Workspace w = carve(workspace, h, B, L1, L2);
        if (w.bytes > workspace_bytes)
            return fail(OETR_E_NOMEM, "oetr_forward: workspace %zu B < required %zu B", workspace_bytes, w.bytes);
        int rc = run_batch(h, w, feat1, feat2, B, a, boxes1, boxes2, true, s, lc);
        if (rc) return rc;
    } else {
        if (nc == 1) sizes[0] = B;
        const size_t need = nc == 1 ? carve(nullptr, h, B, L1, L2).bytes : chunked_bytes(h, B, L1, L2);
        if (need > workspace_bytes)
            return fail(OETR_E_NOMEM, "oetr_forward: workspace %zu B < required %zu B", workspace_bytes, need);
        CU(cudaEventRecord(es.fork, s));
        char* base = static_cast<char*>(workspace);
        int b0 = 0;
        for (int c = 0; c < nc; ++c) {
            const int Bc = sizes[c];
            cudaStream_t sc = es.streams[c];
            CU(cudaStreamWaitEvent(sc, es.fork, 0));
            FwdArgs ac = a;                                  // the sub-batch's slice of the masks
            if (a.mask1) { ac.mask1 = a.mask1 + (size_t)b0 * L1; ac.mask2 = a.mask2 + (size_t)b0 * L2; }
            const float* f1c = feat1 + (size_t)b0 * C * L1;
            const float* f2c = feat2 + (size_t)b0 * C * L2;
            float* b1c = boxes1 + (size_t)b0 * 4;
            float* b2c = boxes2 + (size_t)b0 * 4;
            if (hio) {
                CU(cudaMemcpyAsync(const_cast<float*>(f1c), hio->feat1 + (size_t)b0 * C * L1, (size_t)Bc * L1 * C * sizeof(float),
                                   cudaMemcpyHostToDevice, sc));
                CU(cudaMemcpyAsync(const_cast<float*>(f2c), hio->feat2 + (size_t)b0 * C * L2, (size_t)Bc * L2 * C * sizeof(float),
                                   cudaMemcpyHostToDevice, sc));
            }
            // host requests replay a captured graph of the sub-batch's kernels from its third use on (ChunkGraph)
            static const bool no_graph = getenv("OETR_TIMING") || getenv("OETR_NO_GRAPH");
            ChunkGraph* cg = (hio && hio->graphs && h->prec == OETR_PREC_FP16 && !h->prof.on && !no_graph) ? &hio->graphs[c] : nullptr;
            bool replayed = false;
            if (cg) {
                const int key[11] = {Bc, a.hf1, a.wf1, a.hf2, a.wf2, a.img_h1, a.img_w1, a.img_h2, a.img_w2, a.clamp, nc};
                const void* ptr[7] = {f1c, f2c, b1c, b2c, base, post[0], post[1]};
                const bool same = memcmp(key, cg->key, sizeof(key)) == 0 && memcmp(ptr, cg->ptr, sizeof(ptr)) == 0;
                if (!same) {
                    if (cg->exec) { cudaGraphExecDestroy(cg->exec); cg->exec = nullptr; }
                    memcpy(cg->key, key, sizeof(key)); memcpy(cg->ptr, ptr, sizeof(ptr));
                    cg->uses = 0;
                }
                if (!cg->exec && cg->uses >= 1) {          // second use of this problem: capture (the first ran eagerly)
                    LaunchCounter lg;
                    cudaGraph_t graph = nullptr;
                    CU(cudaStreamBeginCapture(sc, cudaStreamCaptureModeThreadLocal));
                    const Workspace wg = carve(base, h, Bc, L1, L2);
                    const int rcg = run_batch(h, wg, f1c, f2c, Bc, ac, b1c, b2c, false, sc, lg);
                    const cudaError_t ec = cudaStreamEndCapture(sc, &graph);
                    if (rcg == OETR_OK && ec == cudaSuccess && graph &&
                        cudaGraphInstantiate(&cg->exec, graph, 0) == cudaSuccess) cg->launches = lg.n;
                    else { cg->exec = nullptr; cudaGetLastError(); }
                    if (graph) cudaGraphDestroy(graph);
                }
                ++cg->uses;
                if (cg->exec) {
                    CU(cudaGraphLaunch(cg->exec, sc));
                    lc.n += cg->launches;
                    replayed = true;
                }
            }
            const Workspace w = carve(base, h, Bc, L1, L2);
            base += (w.bytes + 1023) & ~size_t(1023);
            if (!replayed) {
                int rc = run_batch(h, w, f1c, f2c, Bc, ac, b1c, b2c, false, sc, lc);
                if (rc) return rc;
            }
            if (hio) {
                CU(cudaMemcpyAsync(hio->boxes1 + (size_t)b0 * 4, b1c, (size_t)Bc * 4 * sizeof(float), cudaMemcpyDeviceToHost, sc));
                CU(cudaMemcpyAsync(hio->boxes2 + (size_t)b0 * 4, b2c, (size_t)Bc * 4 * sizeof(float), cudaMemcpyDeviceToHost, sc));
                CU(cudaMemcpyAsync(hio->flags + c, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, sc));
            }
            CU(cudaEventRecord(es.join[c], sc));
            if (es.join_caller) CU(cudaStreamWaitEvent(s, es.join[c], 0));
            b0 += Bc;
        }
        if (es.n_join) *es.n_join = nc;
    }
    h->last_launches = lc.n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(OETR_E_CUDA, "oetr_forward: launch failed: %s", cudaGetErrorString(e));
    return OETR_OK;
}
}  // namespace

int oetr_forward(oetr_handle* h, const float* feat1, const float* feat2, int batch, int hf1, int wf1, int hf2,
                 int wf2, int img_h1, int img_w1, int img_h2, int img_w2, int clamp, float* boxes1, float* boxes2,
                 float* dbg_hs, float* dbg_memory, float* dbg_cxy, float* dbg_tlbr, void* workspace,
                 size_t workspace_bytes, void* stream) {
    return oetr_forward_masked(h, feat1, feat2, nullptr, nullptr, batch, hf1, wf1, hf2, wf2, img_h1, img_w1, img_h2,
                               img_w2, clamp, boxes1, boxes2, dbg_hs, dbg_memory, dbg_cxy, dbg_tlbr, workspace,
                               workspace_bytes, stream);
}

int oetr_forward_masked(oetr_handle* h, const float* feat1, const float* feat2, const float* mask1, const float* mask2,
                        int batch, int hf1, int wf1, int hf2, int wf2, int img_h1, int img_w1, int img_h2, int img_w2,
                        int clamp, float* boxes1, float* boxes2, float* dbg_hs, float* dbg_memory, float* dbg_cxy,
                        float* dbg_tlbr, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_shapes(h, batch, hf1, wf1, hf2, wf2);
    if (rc) return rc;
    if (batch == 0) { h->last_launches = 0; return OETR_OK; }              // empty batch: nothing to do
    if (!feat1 || !feat2 || !boxes1 || !boxes2 || !workspace)
        return fail(OETR_E_ARG, "oetr_forward: null buffer");
    if (img_h1 < hf1 || img_h2 < hf2 || img_w1 < 1 || img_w2 < 1)
        return fail(OETR_E_SHAPE, "oetr_forward: image sizes (%d,%d),(%d,%d) smaller than the feature maps", img_h1,
                    img_w1, img_h2, img_w2);
    if (reinterpret_cast<uintptr_t>(workspace) & 255)
        return fail(OETR_E_ARG, "oetr_forward: workspace must be 256-byte aligned");
    if ((mask1 == nullptr) != (mask2 == nullptr))
        return fail(OETR_E_ARG, "oetr_forward_masked: give both masks or neither (reference src/model.py:167-171)");
    if (mask1 && h->attn_mode != OETR_ATTN_LINEAR)
        return fail(OETR_E_ARG, "oetr_forward_masked: masks are implemented for linear attention only");
    const FwdArgs a{hf1, wf1, hf2, wf2, img_h1, img_w1, img_h2, img_w2, clamp, dbg_hs, dbg_memory, dbg_cxy, dbg_tlbr,
                    mask1, mask2};
    std::lock_guard<std::mutex> lock(h->mu);
    const EventSet es{h->ev_fork, h->ev_join, true, nullptr, h->aux};
    return forward_core(h, feat1, feat2, batch, a, boxes1, boxes2, workspace, workspace_bytes,
                        static_cast<cudaStream_t>(stream), nullptr, es);
}

int oetr_head_forward(oetr_handle* h, const float* memory1, const float* memory2, const float* hs1, const float* hs2,
                      const float* mask1, const float* mask2, int batch, int hf1, int wf1, int hf2, int wf2, int img_h1,
                      int img_w1, int img_h2, int img_w2, int clamp, float* boxes1, float* boxes2, float* cxy,
                      float* tlbr, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_shapes(h, batch, hf1, wf1, hf2, wf2);
    if (rc) return rc;
    if (batch == 0) return OETR_OK;
    if (!memory1 || !memory2 || !hs1 || !hs2 || !boxes1 || !boxes2 || !workspace)
        return fail(OETR_E_ARG, "oetr_head_forward: null buffer");
    if ((mask1 == nullptr) != (mask2 == nullptr)) return fail(OETR_E_ARG, "oetr_head_forward: give both masks or neither");
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(OETR_E_ARG, "oetr_head_forward: workspace must be 256-byte aligned");
    const int B = batch, L1 = hf1 * wf1, L2 = hf2 * wf2, R1 = B * L1, R2 = B * L2;
    const Workspace w = carve(workspace, h, B, L1, L2);
    if (w.bytes > workspace_bytes) return fail(OETR_E_NOMEM, "oetr_head_forward: workspace %zu B < required %zu B", workspace_bytes, w.bytes);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    LaunchCounter lc;
    const float* W = h->d_w;
    // heat = memory * <memory, hs>; conv3x3 (+bias) -> Y; GroupNorm/ReLU/1x1/softmax/soft-argmax + tlbr + box assembly:
    // the fp32 CUDA-core kernels of either precision's handle (this entry serves the stage-wise Python signatures)
    float *G = w.T, *Gs = w.Q, *Y = w.O;
    heat_scale(memory1, hs1, G, R1, L1, s, lc);
    heat_scale(memory2, hs2, G + (size_t)R1 * C, R2, L2, s, lc);
    head_conv_fp32(h, w, B, hf1, wf1, hf2, wf2, G, Gs, Y, s, lc);
    HeadParams p{};
    p.gn_g = W + h->L.hm_gn_g; p.gn_b = W + h->L.hm_gn_b; p.w3 = W + h->L.hm_w3; p.b3 = W + h->L.hm_b3;
    p.tl_w0 = W + h->L.tl_w0; p.tl_w2 = W + h->L.tl_w2; p.tl_b2 = W + h->L.tl_b2;
    p.batch = B; p.clamp = clamp;
    p.Y = Y; p.hs = hs1; p.hf = hf1; p.wf = wf1; p.img_h = img_h1; p.img_w = img_w1;
    p.boxes = boxes1; p.dbg_cxy = cxy; p.dbg_tlbr = tlbr; p.mask = mask1;
    head_finalize(p, s, lc);
    p.Y = Y + (size_t)R1 * C; p.hs = hs2; p.hf = hf2; p.wf = wf2; p.img_h = img_h2; p.img_w = img_w2;
    p.boxes = boxes2; p.dbg_cxy = cxy ? cxy + 2 * B : nullptr; p.dbg_tlbr = tlbr ? tlbr + 4 * B : nullptr; p.mask = mask2;
    head_finalize(p, s, lc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(OETR_E_CUDA, "oetr_head_forward: launch failed: %s", cudaGetErrorString(e));
    return OETR_OK;
}

static int grow(float** p, size_t* have, size_t need) {
    if (*have >= need) return 0;
    cudaFree(*p); *p = nullptr; *have = 0;
    if (cudaMalloc(p, need * sizeof(float)) != cudaSuccess) return -1;
    *have = need;
    return 0;
}

int oetr_forward_host_submit(oetr_handle* h, const float* feat1_host, const float* feat2_host, int batch, int hf1,
                             int wf1, int hf2, int wf2, int img_h1, int img_w1, int img_h2, int img_w2, int clamp,
                             void* stream, int* ticket) {
    int rc = check_shapes(h, batch, hf1, wf1, hf2, wf2);
    if (rc) return rc;
    if (!ticket) return fail(OETR_E_ARG, "oetr_forward_host_submit: null ticket");
    if (batch > 0 && (!feat1_host || !feat2_host)) return fail(OETR_E_ARG, "oetr_forward_host_submit: null buffer");
    if (batch > 0 && (img_h1 < hf1 || img_h2 < hf2 || img_w1 < 1 || img_w2 < 1))
        return fail(OETR_E_SHAPE, "oetr_forward_host_submit: image sizes (%d,%d),(%d,%d) smaller than the feature maps",
                    img_h1, img_w1, img_h2, img_w2);
    std::lock_guard<std::mutex> lock(h->mu);
    HostSlot& sl = h->slot[h->next_ticket % HOST_SLOTS];
    if (sl.busy)
        return fail(OETR_E_ARG, "oetr_forward_host_submit: %d requests already in flight (wait for ticket %d first)",
                    HOST_SLOTS, sl.ticket);
    sl.batch = batch; sl.n_join = 0;
    if (batch > 0) {
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t n1 = (size_t)batch * C * hf1 * wf1, n2 = (size_t)batch * C * hf2 * wf2;
        size_t ws = 0;
        rc = oetr_workspace_bytes(h, batch, hf1, wf1, hf2, wf2, &ws);
        if (rc) return rc;
        if (grow(&sl.feat1, &sl.feat1_n, n1) || grow(&sl.feat2, &sl.feat2_n, n2) || grow(&sl.boxes, &sl.boxes_n, (size_t)batch * 8))
            return fail(OETR_E_NOMEM, "oetr_forward_host_submit: staging allocation failed");
        if (sl.ws_n < ws) {
            cudaFree(sl.ws); sl.ws = nullptr; sl.ws_n = 0;
            if (cudaMalloc(&sl.ws, ws) != cudaSuccess) return fail(OETR_E_NOMEM, "oetr_forward_host_submit: workspace allocation failed");
            sl.ws_n = ws;
        }
        if (sl.boxes_pin_n < (size_t)batch * 8) {
            cudaFreeHost(sl.boxes_pin); sl.boxes_pin = nullptr; sl.boxes_pin_n = 0;
            if (cudaMallocHost(&sl.boxes_pin, (size_t)batch * 8 * sizeof(float)) != cudaSuccess)
                return fail(OETR_E_NOMEM, "oetr_forward_host_submit: pinned staging allocation failed");
            sl.boxes_pin_n = (size_t)batch * 8;
        }
        const FwdArgs a{hf1, wf1, hf2, wf2, img_h1, img_w1, img_h2, img_w2, clamp, nullptr, nullptr, nullptr, nullptr,
                        nullptr, nullptr};
        const HostIO hio{feat1_host, feat2_host, sl.boxes_pin, sl.boxes_pin + (size_t)batch * 4, sl.flag_pin, sl.cg};
        const EventSet es{sl.ev_fork, sl.ev_join, false, &sl.n_join, sl.aux};
        rc = forward_core(h, sl.feat1, sl.feat2, batch, a, sl.boxes, sl.boxes + (size_t)batch * 4, sl.ws, sl.ws_n, s, &hio, es);
        if (rc) {   // some sub-batches may have been queued: drain them before the slot can be reused
            cudaDeviceSynchronize();
            return rc;
        }
    }
    sl.busy = true;
    sl.ticket = (int)(h->next_ticket & 0x7fffffff);
    *ticket = sl.ticket;
    ++h->next_ticket;
    return OETR_OK;
}

int oetr_forward_host_wait(oetr_handle* h, int ticket, float* boxes1_host, float* boxes2_host) {
    if (!h) return fail(OETR_E_ARG, "null handle");
    HostSlot* sl = nullptr;
    {
        std::lock_guard<std::mutex> lock(h->mu);
        for (HostSlot& c : h->slot) if (c.busy && c.ticket == ticket) sl = &c;
    }
    if (!sl) return fail(OETR_E_ARG, "oetr_forward_host_wait: ticket %d is not in flight", ticket);
    if (sl->batch > 0 && (!boxes1_host || !boxes2_host)) return fail(OETR_E_ARG, "oetr_forward_host_wait: null buffer");
    int flag = 0;
    cudaError_t err = cudaSuccess;
    for (int c = 0; c < sl->n_join; ++c) {
        const cudaError_t e = cudaEventSynchronize(sl->ev_join[c]);
        if (e != cudaSuccess) err = e;
        flag |= sl->flag_pin[c];
    }
    const int batch = sl->batch;
    if (err == cudaSuccess && !flag && batch > 0) {
        memcpy(boxes1_host, sl->boxes_pin, (size_t)batch * 4 * sizeof(float));
        memcpy(boxes2_host, sl->boxes_pin + (size_t)batch * 4, (size_t)batch * 4 * sizeof(float));
    }
    {
        std::lock_guard<std::mutex> lock(h->mu);
        sl->busy = false;
    }
    if (err != cudaSuccess) return fail(OETR_E_CUDA, "oetr_forward_host_wait: %s", cudaGetErrorString(err));
    if (flag) {
        cudaMemset(h->d_flag, 0, sizeof(int));
        return fail(OETR_E_CUDA, "a device-side mbarrier wait timed out (pipeline protocol error); results are invalid");
    }
    return OETR_OK;
}

int oetr_forward_host(oetr_handle* h, const float* feat1_host, const float* feat2_host, int batch, int hf1, int wf1,
                      int hf2, int wf2, int img_h1, int img_w1, int img_h2, int img_w2, int clamp,
                      float* boxes1_host, float* boxes2_host, void* stream) {
    if (batch > 0 && (!boxes1_host || !boxes2_host)) return fail(OETR_E_ARG, "oetr_forward_host: null buffer");
    int ticket = -1;
    int rc = oetr_forward_host_submit(h, feat1_host, feat2_host, batch, hf1, wf1, hf2, wf2, img_h1, img_w1, img_h2,
                                      img_w2, clamp, stream, &ticket);
    if (rc) return rc;
    return oetr_forward_host_wait(h, ticket, boxes1_host, boxes2_host);
}

int oetr_debug_cycles(unsigned long long* out, int n, int reset) {
    if (n < 0) { tc_debug_enable(reset != 0); return 0; }          // n < 0: switch the accumulators on / off
    if (!out || n < 1) return fail(OETR_E_ARG, "oetr_debug_cycles: bad arguments");
    return tc_debug_read(out, n, reset);
}

int oetr_selftest_geometry(int batch, int hf1, int wf1, int hf2, int wf2, int* flat_tiles) {
    if (batch < 1 || hf1 < 1 || wf1 < 1 || hf2 < 1 || wf2 < 1) return fail(OETR_E_ARG, "oetr_selftest_geometry: bad arguments");
    char msg[256] = "";
    const int rc = tc_check_geometry(batch, hf1 * wf1, hf2 * wf2, msg, sizeof(msg));
    if (rc == -1) return fail(OETR_E_SHAPE, "oetr_selftest_geometry: %s", msg);
    if (flat_tiles) *flat_tiles = rc > 0 ? rc : 0;
    return OETR_OK;
}

int oetr_selftest_tcgen05(float* errs_host, int n_errs) {
    if (!errs_host || n_errs < 1) return fail(OETR_E_ARG, "oetr_selftest_tcgen05: bad arguments");
    int dev = 0;
    CU(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(OETR_E_ARCH, "oetr_selftest_tcgen05: device is sm_%d%d", prop.major, prop.minor);
    char msg[256] = "";
    if (tc_selftest(errs_host, n_errs, msg, sizeof(msg)) != 0) return fail(OETR_E_CUDA, "oetr_selftest_tcgen05: %s", msg);
    return OETR_OK;
}

}  // extern "C"
