// Micro-benchmark (debugging aid): per-SM global->shared streaming rate of cp.async.bulk through an mbarrier ring,
// no compute.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bulk_bw tools/bulk_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../imagematching-oetr_b200/csrc/tc_common.cuh"
using namespace oetr::tc;

template <int STAGES>
__global__ void k_stream(const uint8_t* src, size_t src_bytes, uint32_t copy_bytes, int ncopies, int shared_src, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    if (threadIdx.x == 0) { for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_mbar_init(); }
    __syncthreads();
    const uint8_t* base = shared_src ? src : src + ((size_t)blockIdx.x * 2654435761u % (src_bytes / 2 / copy_bytes)) * copy_bytes;
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (int g = 0; g < ncopies; ++g) {
            const int st = g % STAGES;
            mbar_wait(&empty[st], ((g / STAGES) & 1) ^ 1, nullptr);
            mbar_arrive_expect_tx(&full[st], copy_bytes);
            bulk_g2s(smem + (size_t)st * copy_bytes, base + ((size_t)g * copy_bytes) % (src_bytes / 2), copy_bytes, &full[st]);
        }
        (void)t0;
    } else if (threadIdx.x == 32) {
        long long t0 = clock64();
        for (int g = 0; g < ncopies; ++g) {
            const int st = g % STAGES;
            mbar_wait(&full[st], (g / STAGES) & 1, nullptr);
            mbar_arrive(&empty[st]);
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

template <int STAGES>
void run(const uint8_t* d, size_t bytes, uint32_t copy_bytes, int grid, int shared_src, long long* d_cyc) {
    const int ncopies = 512;
    cudaFuncSetAttribute(k_stream<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * copy_bytes);
    for (int rep = 0; rep < 2; ++rep) k_stream<STAGES><<<grid, 64, STAGES * copy_bytes>>>(d, bytes, copy_bytes, ncopies, shared_src, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    printf("stages %d x %5u B, grid %3d, %s source: %.1f B/cycle/SM (%s)\n", STAGES, copy_bytes, grid,
           shared_src ? "same" : "distinct", (double)ncopies * copy_bytes / avg, cudaGetErrorString(e));
}

int main() {
    const size_t bytes = 64u << 20;
    uint8_t* d; long long* d_cyc;
    cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes); cudaMalloc(&d_cyc, 148 * sizeof(long long));
    for (int grid : {1, 16, 148}) for (int shared : {1, 0}) {
        run<2>(d, bytes, 16384, grid, shared, d_cyc);
        run<4>(d, bytes, 16384, grid, shared, d_cyc);
        run<6>(d, bytes, 16384, grid, shared, d_cyc);
        run<12>(d, bytes, 16384, grid, shared, d_cyc);
        run<6>(d, bytes, 32768, grid, shared, d_cyc);
        run<24>(d, bytes, 4096, grid, shared, d_cyc);
    }
    return 0;
}
