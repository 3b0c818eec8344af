// Micro-benchmark: issue cost of packed fp32 (fma.rn.f32x2 -> SASS FFMA2) vs scalar FFMA on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu ; run on the B200.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a); rb = *reinterpret_cast<unsigned long long*>(&b);
    rc = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

template <int MODE>  // 0: scalar FFMA (8 independent chains), 1: FFMA2 (8 independent packed chains = 16 lanes of work)
__global__ void k(float* out, int iters, float s) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 m = make_float2(s, s * 0.999f), c = make_float2(1e-3f, 2e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); }
            else a[i] = ffma2(a[i], m, c);
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.9999f); else k<1><<<148 * 8, 256>>>(out, iters, 0.9999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double inst = 148.0 * 8 * 8 * 8.0 * iters;               // warp-instructions of the FMA kind
            const double flops = inst * 32 * 2 * (mode ? 2 : 1);
            if (rep) printf("mode %d (%s): %.3f ms  %.1f G warp-inst/s  %.1f TFLOP/s\n", mode, mode ? "FFMA2" : "FFMA", ms, inst / ms * 1e-6, flops / ms * 1e-9);
        }
    }
    return 0;
}
