// corr.cu -- 1-D stereo correlation lookup (replaces the `corr_sampler`
// extension the reference calls at core/corr.py:22 (forward) and :28 (backward)).
// Arithmetic in fp32 for both volume dtypes; fp16 volumes are converted on load/store.
#include <cuda_fp16.h>
#include "gpsg_internal.cuh"

namespace gpsg {

template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// thread = (n, y, x); reads 2r+2 consecutive taps of its volume row, writes 2r+1 outputs (coalesced over x).
template <typename T, int R>
__global__ void __launch_bounds__(256) corr_fwd_kernel(int B, int H, int W1, int W2, const T* __restrict__ vol,
                                                       int64_t sb, int64_t sh, int64_t sw1,
                                                       const float* __restrict__ coords, int64_t csb, int r_dyn,
                                                       T* __restrict__ out) {
    const int r = R > 0 ? R : r_dyn;
    const int rd = 2 * r + 1;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)B * H * W1;
    if (idx >= total) return;
    const int x = (int)(idx % W1);
    const int y = (int)((idx / W1) % H);
    const int n = (int)(idx / ((int64_t)W1 * H));
    const float x0 = coords[n * csb + (int64_t)y * W1 + x];
    const float fl = floorf(x0);
    const float dx = x0 - fl;
    const int xf = (int)fl;
    const T* row = vol + n * sb + y * sh + x * sw1;
    float prev = 0.f;  // tap i-1
    {
        const int x1 = xf - r;
        if (x1 >= 0 && x1 < W2) prev = ld_f<T>(row + x1);
    }
    const int64_t plane = (int64_t)H * W1;
    T* o = out + (int64_t)n * rd * plane + (int64_t)y * W1 + x;
#pragma unroll
    for (int i = 1; i <= rd; ++i) {  // rd is a compile-time constant when R > 0
        const int x1 = xf - r + i;
        float cur = 0.f;
        if (x1 >= 0 && x1 < W2) cur = ld_f<T>(row + x1);
        // out[i-1] = tap(i-1)*(1-dx) + tap(i)*dx
        st_f<T>(o + (int64_t)(i - 1) * plane, prev * (1.0f - dx) + cur * dx);
        prev = cur;
    }
}

// thread = 4 consecutive x1 of one (n,y,x) row of grad_volume (fully written, zeros outside the window)
template <typename T>
__global__ void __launch_bounds__(256) corr_bwd_kernel(int B, int H, int W1, int W2, const float* __restrict__ coords,
                                                       int64_t csb, const T* __restrict__ gout, int r,
                                                       T* __restrict__ gvol) {
    const int rd = 2 * r + 1;
    const int chunks = (W2 + 3) / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)B * H * W1 * chunks;
    if (idx >= total) return;
    const int ch = (int)(idx % chunks);
    const int64_t rowi = idx / chunks;
    const int x = (int)(rowi % W1);
    const int y = (int)((rowi / W1) % H);
    const int n = (int)(rowi / ((int64_t)W1 * H));
    const float x0 = coords[n * csb + (int64_t)y * W1 + x];
    const float fl = floorf(x0);
    const float dx = x0 - fl;
    const int xf = (int)fl;
    const int64_t plane = (int64_t)H * W1;
    const T* go = gout + (int64_t)n * rd * plane + (int64_t)y * W1 + x;
    T* row = gvol + rowi * W2;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int x1 = ch * 4 + e;
        if (x1 >= W2) break;
        const int i = x1 - xf + r;
        float g = 0.f;
        if (i >= 0 && i <= rd) {
            if (i > 0) g += ld_f<T>(go + (int64_t)(i - 1) * plane) * dx;
            if (i < rd) g += ld_f<T>(go + (int64_t)i * plane) * (1.0f - dx);
        }
        st_f<T>(row + x1, g);
    }
}

int launch_corr_fwd(int dtype, int B, int H, int W1, int W2, const void* vol, int64_t sb, int64_t sh, int64_t sw1,
                    const float* coords, int64_t csb, int r, void* out, cudaStream_t stream) {
    const int64_t total = (int64_t)B * H * W1;
    if (total == 0) return GPSG_OK;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (dtype == 0) {
        if (r == 4) corr_fwd_kernel<float, 4><<<grid, 256, 0, stream>>>(B, H, W1, W2, (const float*)vol, sb, sh, sw1, coords, csb, r, (float*)out);
        else corr_fwd_kernel<float, 0><<<grid, 256, 0, stream>>>(B, H, W1, W2, (const float*)vol, sb, sh, sw1, coords, csb, r, (float*)out);
    } else {
        if (r == 4) corr_fwd_kernel<__half, 4><<<grid, 256, 0, stream>>>(B, H, W1, W2, (const __half*)vol, sb, sh, sw1, coords, csb, r, (__half*)out);
        else corr_fwd_kernel<__half, 0><<<grid, 256, 0, stream>>>(B, H, W1, W2, (const __half*)vol, sb, sh, sw1, coords, csb, r, (__half*)out);
    }
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

int launch_corr_bwd(int dtype, int B, int H, int W1, int W2, const float* coords, int64_t csb, const void* gout, int r,
                    void* gvol, cudaStream_t stream) {
    const int64_t total = (int64_t)B * H * W1 * ((W2 + 3) / 4);
    if (total == 0) return GPSG_OK;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (dtype == 0) corr_bwd_kernel<float><<<grid, 256, 0, stream>>>(B, H, W1, W2, coords, csb, (const float*)gout, r, (float*)gvol);
    else corr_bwd_kernel<__half><<<grid, 256, 0, stream>>>(B, H, W1, W2, coords, csb, (const __half*)gout, r, (__half*)gvol);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

}  // namespace gpsg
