// Micro-benchmark (debugging aid): issue rate of the instructions the row-warp epilogues are made of, 16 warps
// per SM (4 per scheduler) like k_enc's row warps.  Prints cycles per warp-instruction per scheduler.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/pipe_rate tools/pipe_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define U64(x) reinterpret_cast<unsigned long long&>(x)
constexpr int ITERS = 2048, NACC = 8;

template <int OP>
__global__ void k(float* out, float seed, long long* cyc) {
    float2 a[NACC];
    uint32_t u[NACC];
    for (int i = 0; i < NACC; ++i) { a[i] = make_float2(seed + i + threadIdx.x, seed * 0.5f + i); u[i] = __float_as_uint(a[i].x); }
    float2 b = make_float2(seed * 0.999f, seed * 1.001f), c = make_float2(0.25f * seed, 0.125f * seed);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (OP == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); }                                        // FFMA (3-reg)
            if (OP == 1) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(U64(a[i])) : "l"(U64(b)), "l"(U64(c))); }  // FFMA2
            if (OP == 2) { a[i].x = a[i].x + b.x; }                                                  // FADD
            if (OP == 3) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(U64(a[i])) : "l"(U64(b))); }                 // FADD2
            if (OP == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i].x)); }              // MUFU.EX2
            if (OP == 5) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i].x), "f"(__uint_as_float(u[i]))); }   // F2FP
            if (OP == 6) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(__float_as_uint(b.x)), "r"(__float_as_uint(c.x))); }   // LOP3
            if (OP == 8) { a[i].x = fmaf(a[i].x, 1.0001f, 0.5f); }                                   // FFMA imm
            if (OP == 9) { a[i].x = fmaf(a[i].x, b.x, c.x); u[i] = (u[i] & 0xFFFFE000u) ^ __float_as_uint(c.y); }   // FFMA + LOP3 mix
            if (OP == 10) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(U64(a[i])) : "l"(U64(b)), "l"(U64(c)));
                            u[i] = (u[i] & 0xFFFFE000u) ^ __float_as_uint(c.y); }                    // FFMA2 + LOP3 mix
            if (OP == 11) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i].x));
                            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(U64(a[(i + 1) % NACC])) : "l"(U64(b)), "l"(U64(c))); }  // MUFU + FFMA2
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter, float* d_out, long long* d_cyc) {
    k<OP><<<148, 512>>>(d_out, 1.0f, d_cyc);
    k<OP><<<148, 512>>>(d_out, 1.0f, d_cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    // 4 warps per scheduler, each ITERS * NACC * per_iter instructions
    printf("%-28s %.2f cycles per warp-instruction per scheduler\n", name, avg / (4.0 * ITERS * NACC * per_iter));
}

int main() {
    float* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 148 * 512 * 4); cudaMalloc(&d_cyc, 148 * 8);
    run<0>("FFMA (3 registers)", 1, d_out, d_cyc);
    run<8>("FFMA (immediates)", 1, d_out, d_cyc);
    run<1>("FFMA2", 1, d_out, d_cyc);
    run<2>("FADD", 1, d_out, d_cyc);
    run<3>("FADD2", 1, d_out, d_cyc);
    run<4>("MUFU.EX2", 1, d_out, d_cyc);
    run<5>("F2FP.F16.F32.PACK_AB", 1, d_out, d_cyc);
    run<6>("LOP3", 1, d_out, d_cyc);
    run<9>("FFMA + LOP3 (per instr)", 2, d_out, d_cyc);
    run<10>("FFMA2 + LOP3 (per instr)", 2, d_out, d_cyc);
    run<11>("MUFU + FFMA2 (per instr)", 2, d_out, d_cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
