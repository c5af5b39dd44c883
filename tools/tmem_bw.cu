// Micro-benchmark (debugging aid): TMEM -> register (tcgen05.ld) and register -> TMEM (tcgen05.st) throughput of one
// SM with 4..16 warps, and the rate of epilogue-style code around it.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/tmem_bw tools/tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../imagematching-oetr_b200/csrc/tc_common.cuh"
using namespace oetr::tc;

// mode 0: ld 32x32b.x32 + wait per load; 1: two loads back to back then one wait; 2: st 32x32b.x32; 3: ld.x32 + 32 FADD
template <int mode>
__global__ void __launch_bounds__(576, 1) k_tmem(int nwarps, int iters, long long* cycles, float* sink) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 16) tmem_alloc(&tbase, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tbase;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = (float)lane;
    if (warp < nwarps) {
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t col = (warp >> 2) * 32;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t a = tm + lane_addr + ((col + (it & 1) * 128 + ((it >> 1) & 1) * 256) & 511);
            if constexpr (mode == 0 || mode == 3) {
                float v[32];
                tmem_ld32(a, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] += v[i];
            } else if constexpr (mode == 1) {
                uint32_t r[64];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                    "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n\t"
                    "tcgen05.wait::ld.sync.aligned;"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                      "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                      "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                      "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
                      "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
                      "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                    : "r"(tm + lane_addr + ((it & 1) * 256) + (warp >> 2) * 64)
                    : "memory");
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(r[i]) + __uint_as_float(r[i + 32]);
            } else if constexpr (mode == 2) {
                tmem_st32(a, acc);
                tmem_st_wait();
            } else if constexpr (mode == 4) {          // two x32 loads, ONE wait
                uint32_t r[64];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%64];\n\t"
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%65];\n\t"
                    "tcgen05.wait::ld.sync.aligned;"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                      "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                      "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                      "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
                      "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
                      "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                    : "r"(a), "r"(a ^ 128)
                    : "memory");
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(r[i]) + __uint_as_float(r[i + 32]);
            } else if constexpr (mode == 5) {          // x32 load + wait, then ~256 dependent-free FFMA per thread (epilogue-like math)
                float v[32];
                tmem_ld32(a, v);
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], 1.0001f, v[i]);
            } else if constexpr (mode == 6) {          // math only (the same 256 FFMA)
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], 1.0001f, 0.5f);
            }
        }
        const long long t1 = clock64();
        if (lane == 0) cycles[warp] = t1 - t0;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    sink[threadIdx.x] = s;
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tm, 512);
}

int main() {
    long long* d_cyc; float* d_sink;
    cudaMalloc(&d_cyc, 32 * sizeof(long long)); cudaMalloc(&d_sink, 576 * sizeof(float));
    const int iters = 256;
    const char* names[] = {"ld 32x32b.x32 (4 KB/warp) + wait", "ld 32x32b.x64 (8 KB/warp) + wait", "st 32x32b.x32 + wait", "",
                           "2 x ld.x32, one wait (8 KB/warp)", "ld.x32 + wait + 256 FFMA", "256 FFMA only"};
    for (int mode : {0, 1, 2, 4, 5, 6})
        for (int nw : {1, 4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                switch (mode) {
                    case 0: k_tmem<0><<<1, 576>>>(nw, iters, d_cyc, d_sink); break;
                    case 1: k_tmem<1><<<1, 576>>>(nw, iters, d_cyc, d_sink); break;
                    case 2: k_tmem<2><<<1, 576>>>(nw, iters, d_cyc, d_sink); break;
                    case 4: k_tmem<4><<<1, 576>>>(nw, iters, d_cyc, d_sink); break;
                    case 5: k_tmem<5><<<1, 576>>>(nw, iters, d_cyc, d_sink); break;
                    default: k_tmem<6><<<1, 576>>>(nw, iters, d_cyc, d_sink); break;
                }
            }
            cudaError_t e = cudaDeviceSynchronize();
            long long h[32];
            cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
            const double bytes = (double)nw * iters * ((mode == 1 || mode == 4) ? 8192 : 4096);
            printf("%-36s %2d warps: %7.1f cycles per instruction per warp, %6.1f B/cycle/SM (%s)\n", names[mode], nw,
                   (double)mx / iters, bytes / mx, cudaGetErrorString(e));
        }
    return 0;
}
