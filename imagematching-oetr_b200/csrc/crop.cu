// Post-box plumbing on the device: the crop + bicubic resize the reference does on the host for
// every pair -- dloc/core/utils/utils.py:510-564 `tensor_overlap_crop`: slice the box out of the image tensor, D2H,
// * 255, cv2.resize(float32, INTER_CUBIC) (once, or twice when the matcher needs a size divisor), / 255, H2D.  Here the image
// never leaves the GPU: one kernel per resize pass, every job of a batch in the same launch.
// Arithmetic follows cv2's INTER_CUBIC (opencv-python 4.13, optimised build): per axis the source position
// f = (d + 0.5) * src/dst - 0.5, its fraction and the four Keys weights (A = -0.75) in double, weights rounded to float32,
// source indices clamped to the crop, horizontal pass then vertical pass in float32.  Equal sizes are a plain copy.
// HBM-bound gather work: no tensor cores; threads map to output pixels, x fastest (coalesced stores, L1-resident taps).
#include "../../include/oetr_b200.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>

namespace {

constexpr int MAX_JOBS = 32;

struct Job {
    const float* src;   // first element of the FULL image plane 0: [channels][src_h][src_w]
    float* dst;         // [channels][new_h][new_w]
    int channels, src_h, src_w, x0, y0, crop_w, crop_h, new_w, new_h, flags;
    int block_begin;    // first block of this job in the launch
    int tiles_x;
};
struct Jobs { Job j[MAX_JOBS]; int n; };

__device__ __forceinline__ void cubic_weights(int d, int src, int dst, int& s, float (&w)[4]) {
    const double scale = 1.0 / ((double)dst / (double)src);
    const double f = ((double)d + 0.5) * scale - 0.5;
    const double fl = floor(f);
    const double t = f - fl, A = -0.75;
    s = (int)fl;
    const double t1 = t + 1.0, u = 1.0 - t;
    const double w0 = ((A * t1 - 5.0 * A) * t1 + 8.0 * A) * t1 - 4.0 * A;
    const double w1 = ((A + 2.0) * t - (A + 3.0)) * t * t + 1.0;
    const double w2 = ((A + 2.0) * u - (A + 3.0)) * u * u + 1.0;
    w[0] = (float)w0; w[1] = (float)w1; w[2] = (float)w2; w[3] = (float)(1.0 - w0 - w1 - w2);
}

constexpr int TX = 32, TY = 8;

__global__ void __launch_bounds__(TX * TY) k_crop_resize(const __grid_constant__ Jobs jobs) {
    int ji = 0;
#pragma unroll 1
    while (ji + 1 < jobs.n && (int)blockIdx.x >= jobs.j[ji + 1].block_begin) ++ji;
    const Job& J = jobs.j[ji];
    const int b = blockIdx.x - J.block_begin;
    const int x = (b % J.tiles_x) * TX + threadIdx.x, y = (b / J.tiles_x) * TY + threadIdx.y;
    if (x >= J.new_w || y >= J.new_h) return;
    const float in_mul = (J.flags & OETR_CROP_MUL255) ? 255.f : 1.f;
    const bool div = J.flags & OETR_CROP_DIV255;
    const size_t plane = (size_t)J.src_h * J.src_w, oplane = (size_t)J.new_h * J.new_w;
    if (J.new_w == J.crop_w && J.new_h == J.crop_h) {                    // cv2.resize with dsize == ssize copies
        for (int c = 0; c < J.channels; ++c) {
            const float v = J.src[c * plane + (size_t)(J.y0 + y) * J.src_w + J.x0 + x] * in_mul;
            J.dst[c * oplane + (size_t)y * J.new_w + x] = div ? v / 255.f : v;
        }
        return;
    }
    int sx, sy;
    float wx[4], wy[4];
    cubic_weights(x, J.crop_w, J.new_w, sx, wx);
    cubic_weights(y, J.crop_h, J.new_h, sy, wy);
    int ix[4], iy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ix[k] = J.x0 + min(max(sx - 1 + k, 0), J.crop_w - 1);
        iy[k] = J.y0 + min(max(sy - 1 + k, 0), J.crop_h - 1);
    }
    for (int c = 0; c < J.channels; ++c) {
        const float* p = J.src + c * plane;
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const float* row = p + (size_t)iy[ky] * J.src_w;
            float r = __fmul_rn(__ldg(row + ix[0]) * in_mul, wx[0]);           // separate multiplies and adds, like the
            r = __fadd_rn(r, __fmul_rn(__ldg(row + ix[1]) * in_mul, wx[1]));   // two float32 passes of the host code
            r = __fadd_rn(r, __fmul_rn(__ldg(row + ix[2]) * in_mul, wx[2]));
            r = __fadd_rn(r, __fmul_rn(__ldg(row + ix[3]) * in_mul, wx[3]));
            acc = ky == 0 ? __fmul_rn(r, wy[0]) : __fadd_rn(acc, __fmul_rn(r, wy[ky]));
        }
        J.dst[c * oplane + (size_t)y * J.new_w + x] = div ? acc / 255.f : acc;
    }
}

thread_local char g_cerr[256] = "";
int cfail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_cerr, sizeof(g_cerr), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace

extern "C" {

const char* oetr_crop_last_error(void) { return g_cerr; }

int oetr_crop_resize(const oetr_crop_job* jobs, int n_jobs, void* stream) {
    if (n_jobs == 0) return OETR_OK;
    if (!jobs || n_jobs < 0) return cfail(OETR_E_ARG, "oetr_crop_resize: null / negative job list");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int base = 0; base < n_jobs; base += MAX_JOBS) {
        Jobs J;
        J.n = n_jobs - base < MAX_JOBS ? n_jobs - base : MAX_JOBS;
        int blocks = 0;
        for (int i = 0; i < J.n; ++i) {
            const oetr_crop_job& a = jobs[base + i];
            if (!a.src || !a.dst) return cfail(OETR_E_ARG, "oetr_crop_resize: job %d has a null pointer", base + i);
            if (a.channels < 1 || a.src_h < 1 || a.src_w < 1 || a.new_w < 1 || a.new_h < 1 || a.new_w > 16384 || a.new_h > 16384)
                return cfail(OETR_E_SHAPE, "oetr_crop_resize: job %d: channels %d image %d x %d -> %d x %d", base + i, a.channels,
                             a.src_h, a.src_w, a.new_h, a.new_w);
            // Python slicing of image[..., y0:y1, x0:x1]: the end is clipped to the image, an empty slice is an error here
            const int x1 = a.x1 < a.src_w ? a.x1 : a.src_w, y1 = a.y1 < a.src_h ? a.y1 : a.src_h;
            if (a.x0 < 0 || a.y0 < 0 || x1 - a.x0 < 1 || y1 - a.y0 < 1)
                return cfail(OETR_E_SHAPE, "oetr_crop_resize: job %d: empty or negative crop [%d:%d, %d:%d] of %d x %d", base + i,
                             a.y0, a.y1, a.x0, a.x1, a.src_h, a.src_w);
            Job& j = J.j[i];
            j.src = a.src; j.dst = a.dst; j.channels = a.channels; j.src_h = a.src_h; j.src_w = a.src_w;
            j.x0 = a.x0; j.y0 = a.y0; j.crop_w = x1 - a.x0; j.crop_h = y1 - a.y0; j.new_w = a.new_w; j.new_h = a.new_h;
            j.flags = a.flags;
            j.tiles_x = (a.new_w + TX - 1) / TX;
            j.block_begin = blocks;
            blocks += j.tiles_x * ((a.new_h + TY - 1) / TY);
        }
        k_crop_resize<<<blocks, dim3(TX, TY), 0, s>>>(J);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cfail(OETR_E_CUDA, "oetr_crop_resize: %s", cudaGetErrorString(e));
    }
    return OETR_OK;
}

}  // extern "C"
