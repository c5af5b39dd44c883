// oetr_gather_*: the boxes of every rank on every rank WITHOUT a collective ("fused peer-store
// epilogue").  One process per GPU; every rank owns a cudaMalloc'ed communication buffer that its peers map through CUDA
// IPC.  A submit is one tiny kernel that stores this rank's [pairs][2][4] boxes straight into every peer's buffer over
// NVLink (plain st.global to peer memory) and then publishes a per-(slot, source) step flag with a system-scope release;
// a collect is one tiny kernel that acquires the flags of all sources for the oldest outstanding step, copies the slot
// into the caller's output and advances this rank's `consumed` counter, which the writers read remotely for flow control
// (a ring of `slots` steps may be in flight; ranks never rendezvous, they only wait for data they need).
// No NCCL kernel, no host synchronisation; both kernels are stream-ordered.
#include "../../include/oetr_b200.h"

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

namespace {
thread_local char g_gerr[512] = "";
int gfail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_gerr, sizeof(g_gerr), fmt, ap);
    va_end(ap);
    return code;
}
constexpr int MAX_WORLD = 16;
struct Peers { char* base[MAX_WORLD]; };

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// layout of one rank's buffer: data [slots][world][pairs*8] floats | flags [slots][world] u64 | consumed u64
__host__ __device__ inline size_t data_floats(int slots, int world, int pairs) { return (size_t)slots * world * pairs * 8; }
__host__ __device__ inline size_t flags_off(int slots, int world, int pairs) { return (data_floats(slots, world, pairs) * 4 + 255) & ~size_t(255); }
__host__ __device__ inline size_t consumed_off(int slots, int world, int pairs) { return flags_off(slots, world, pairs) + (size_t)slots * world * 8; }

// grid = world CTAs: CTA p pushes this rank's boxes of step `step` into rank p's buffer
__global__ void __launch_bounds__(128) k_gather_push(Peers peers, const float* __restrict__ b1, const float* __restrict__ b2, int world,
                                                     int rank, int pairs, int slots, unsigned long long step) {
    const int p = blockIdx.x;
    char* dst = peers.base[p];
    const int slot = (int)(step % slots);
    // flow control: the slot is free once rank p has collected step - slots
    if (step >= (unsigned long long)slots && threadIdx.x == 0) {
        const unsigned long long* consumed = reinterpret_cast<const unsigned long long*>(dst + consumed_off(slots, world, pairs));
        unsigned long long spins = 0;
        while (ld_acquire_sys(consumed) + slots <= step) {
            if (++spins > (1ull << 26)) __trap();       // a peer that never collects: fail loudly instead of hanging
            __nanosleep(200);
        }
    }
    __syncthreads();
    float* out = reinterpret_cast<float*>(dst) + ((size_t)slot * world + rank) * pairs * 8;
    for (int i = threadIdx.x; i < pairs * 8; i += blockDim.x) {
        const int pr = i >> 3, k = i & 7;                // [pair][image 0/1][4]
        out[i] = k < 4 ? b1[pr * 4 + k] : b2[pr * 4 + k - 4];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        st_release_sys(reinterpret_cast<unsigned long long*>(dst + flags_off(slots, world, pairs)) + (size_t)slot * world + rank, step + 1);
}

// one CTA: waits for the flags of every source for `step`, copies the slot out, releases it
__global__ void __launch_bounds__(128) k_gather_collect(char* mine, float* __restrict__ all_boxes, int world, int pairs, int slots,
                                                        unsigned long long step) {
    const int slot = (int)(step % slots);
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(mine + flags_off(slots, world, pairs)) + (size_t)slot * world;
    if (threadIdx.x < world) {
        unsigned long long spins = 0;
        while (ld_acquire_sys(flags + threadIdx.x) != step + 1) {
            if (++spins > (1ull << 26)) __trap();
            __nanosleep(200);
        }
    }
    __syncthreads();
    const float* src = reinterpret_cast<const float*>(mine) + (size_t)slot * world * pairs * 8;
    for (int i = threadIdx.x; i < world * pairs * 8; i += blockDim.x) all_boxes[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        st_release_sys(reinterpret_cast<unsigned long long*>(mine + consumed_off(slots, world, pairs)), step + 1);
}
}  // namespace

struct oetr_gather {
    int world = 1, rank = 0, pairs = 0, slots = 0, device = 0;
    char* mine = nullptr;
    size_t bytes = 0;
    Peers peers{};
    bool opened[MAX_WORLD] = {};
    bool connected = false;
    unsigned long long pushed = 0, collected = 0;
};

#define GCU(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return gfail(OETR_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_));         \
    } while (0)

extern "C" {

const char* oetr_gather_last_error(void) { return g_gerr; }

int oetr_gather_create(int world, int rank, int pairs_per_rank, int slots, oetr_gather** out, void* ipc_handle_out) {
    if (!out || !ipc_handle_out) return gfail(OETR_E_ARG, "oetr_gather_create: null argument");
    *out = nullptr;
    if (world < 1 || world > MAX_WORLD || rank < 0 || rank >= world || pairs_per_rank < 1 || slots < 1 || slots > 64)
        return gfail(OETR_E_ARG, "oetr_gather_create: world %d rank %d pairs %d slots %d out of range", world, rank, pairs_per_rank, slots);
    oetr_gather* g = new (std::nothrow) oetr_gather();
    if (!g) return gfail(OETR_E_NOMEM, "oetr_gather_create: host allocation failed");
    g->world = world; g->rank = rank; g->pairs = pairs_per_rank; g->slots = slots;
    cudaError_t e = cudaGetDevice(&g->device);
    g->bytes = consumed_off(slots, world, pairs_per_rank) + 256;
    if (e == cudaSuccess) e = cudaMalloc(&g->mine, g->bytes);
    if (e == cudaSuccess) e = cudaMemset(g->mine, 0, g->bytes);
    cudaIpcMemHandle_t hdl;
    memset(&hdl, 0, sizeof(hdl));
    if (e == cudaSuccess && world > 1) e = cudaIpcGetMemHandle(&hdl, g->mine);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(g->mine);
        delete g;
        return gfail(e == cudaErrorMemoryAllocation ? OETR_E_NOMEM : OETR_E_CUDA, "oetr_gather_create: %s", cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == OETR_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(ipc_handle_out, &hdl, sizeof(hdl));
    g->peers.base[rank] = g->mine;
    if (world == 1) g->connected = true;
    *out = g;
    return OETR_OK;
}

int oetr_gather_connect(oetr_gather* g, const void* all_handles) {
    if (!g || !all_handles) return gfail(OETR_E_ARG, "oetr_gather_connect: null argument");
    if (g->connected) return OETR_OK;
    const char* hs = static_cast<const char*>(all_handles);
    for (int p = 0; p < g->world; ++p) {
        if (p == g->rank) continue;
        cudaIpcMemHandle_t hdl;
        memcpy(&hdl, hs + (size_t)p * OETR_IPC_HANDLE_BYTES, sizeof(hdl));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, hdl, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return gfail(OETR_E_CUDA, "oetr_gather_connect: cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
        g->peers.base[p] = static_cast<char*>(ptr);
        g->opened[p] = true;
    }
    g->connected = true;
    return OETR_OK;
}

int oetr_gather_submit(oetr_gather* g, const float* boxes1, const float* boxes2, void* stream) {
    if (!g || !boxes1 || !boxes2) return gfail(OETR_E_ARG, "oetr_gather_submit: null argument");
    if (!g->connected) return gfail(OETR_E_ARG, "oetr_gather_submit: oetr_gather_connect has not been called");
    if (g->pushed - g->collected >= (unsigned long long)g->slots)
        return gfail(OETR_E_ARG, "oetr_gather_submit: %d steps already in flight (collect first)", g->slots);
    k_gather_push<<<g->world, 128, 0, static_cast<cudaStream_t>(stream)>>>(g->peers, boxes1, boxes2, g->world, g->rank, g->pairs, g->slots, g->pushed);
    GCU(cudaGetLastError());
    ++g->pushed;
    return OETR_OK;
}

int oetr_gather_collect(oetr_gather* g, float* all_boxes, void* stream) {
    if (!g || !all_boxes) return gfail(OETR_E_ARG, "oetr_gather_collect: null argument");
    if (g->collected >= g->pushed) return gfail(OETR_E_ARG, "oetr_gather_collect: nothing submitted");
    k_gather_collect<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(g->mine, all_boxes, g->world, g->pairs, g->slots, g->collected);
    GCU(cudaGetLastError());
    ++g->collected;
    return OETR_OK;
}

int oetr_gather_destroy(oetr_gather* g) {
    if (!g) return OETR_OK;
    cudaDeviceSynchronize();
    for (int p = 0; p < g->world; ++p)
        if (g->opened[p]) cudaIpcCloseMemHandle(g->peers.base[p]);
    cudaFree(g->mine);
    delete g;
    return OETR_OK;
}

}  // extern "C"
