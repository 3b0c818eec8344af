// loss.cu -- fused photometric loss on the rasterizer's output: the consumer that turns the
// rendered image into the dL/dimage that gpsg_rasterize_backward takes.
//
// Reference: train_stage2.py:70-72   loss = 0.8 * l1_loss(img, gt) + 0.2 * (1 - ssim(img, gt))
//            lib/loss.py:35-36 (l1_loss), :39-72 (ssim: 11x11 Gaussian window sigma 1.5, depthwise conv2d with zero
//            padding 5, C1 = 0.01^2, C2 = 0.03^2, mean over B*C*H*W).
// The reference runs 5 depthwise 11x11 convolutions + ~15 elementwise kernels forward and autograd's mirror of them
// backward.  Here: ONE forward kernel (separable 11+11 taps in shared memory for the five moments, SSIM map, L1, both
// means, and the three per-pixel partials dS/dmu1, dS/dE[x^2], dS/dE[xy]) and ONE backward kernel (separable
// convolution of those three maps with the same symmetric window + the L1 sign term) -> dL/dimg.
// HBM-bound: forward reads 8 B and writes 12 B per pixel-channel, backward reads 20 B and writes 4 B.
#include "gpsg_internal.cuh"

namespace gpsg {

constexpr int kWin = 11, kHalf = 5;
constexpr int kTile = 32;                        // output tile 32 x 32 per CTA, 256 threads, 4 pixels each
constexpr int kHalo = kTile + 2 * kHalf;         // 42
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

struct Window { float w[kWin]; };

// lib/loss.py:39-41: gauss(x) = exp(-(x - 5)^2 / (2 * 1.5^2)) / sum, as float32.  The bit patterns are the ones torch
// produces for `gaussian(11, 1.5)` (a one-ulp difference in the normalisation is visible in low-variance regions);
// tests/test_loss.py checks them against the window stored with the golden vectors.
static Window make_window() {
    static const uint32_t bits[6] = {981912246u, 1006173953u, 1024685452u, 1038088319u, 1046093343u, 1049113264u};
    Window g;
    for (int i = 0; i < kWin; ++i) {
        const uint32_t b = bits[i <= kHalf ? i : kWin - 1 - i];
        memcpy(&g.w[i], &b, sizeof(float));
    }
    return g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// partial[0][cta] = sum |x - y|, partial[1][cta] = sum ssim; the last CTA (atomic ticket) adds them in index order, so
// the three reported scalars are deterministic.
__global__ void __launch_bounds__(256) l1_ssim_forward_kernel(int planes, int H, int W, const float* __restrict__ img,
                                                              const float* __restrict__ gt, const __grid_constant__ Window win,
                                                              float w_l1, float w_ssim, float* __restrict__ out3,
                                                              float* __restrict__ dmaps, double* __restrict__ partial,
                                                              unsigned* __restrict__ ticket) {
    __shared__ float sx[kHalo][kHalo + 1], sy[kHalo][kHalo + 1];
    __shared__ float hs[5][kHalo][kTile];
    __shared__ double red[2][8];
    __shared__ bool is_last;
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t base = (size_t)plane * H * W;
    const int tid = threadIdx.x;
    for (int e = tid; e < kHalo * kHalo; e += 256) {
        const int r = e / kHalo, c = e % kHalo;
        const int gy = y0 + r - kHalf, gx = x0 + c - kHalf;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        sx[r][c] = in ? img[base + (size_t)gy * W + gx] : 0.f;
        sy[r][c] = in ? gt[base + (size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < kHalo * kTile; e += 256) {       // horizontal pass: 5 moments per (halo row, tile column)
        const int r = e / kTile, c = e % kTile;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float w = win.w[k], x = sx[r][c + k], y = sy[r][c + k];
            a = fmaf(w, x, a); b = fmaf(w, y, b);
            aa = fmaf(w, x * x, aa); bb = fmaf(w, y * y, bb); ab = fmaf(w, x * y, ab);
        }
        hs[0][r][c] = a; hs[1][r][c] = b; hs[2][r][c] = aa; hs[3][r][c] = bb; hs[4][r][c] = ab;
    }
    __syncthreads();
    float l1_acc = 0.f, ss_acc = 0.f;
    const int c = tid & 31;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (tid >> 5) + 8 * i;
        const int gy = y0 + r, gx = x0 + c;
        if (gy >= H || gx >= W) continue;
        float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float w = win.w[k];
#pragma unroll
            for (int q = 0; q < 5; ++q) m[q] = fmaf(w, hs[q][r + k][c], m[q]);
        }
        const float mu1 = m[0], mu2 = m[1];
        const float s11 = m[2] - mu1 * mu1, s22 = m[3] - mu2 * mu2, s12 = m[4] - mu1 * mu2;
        const float A1 = 2.f * mu1 * mu2 + kC1, A2 = 2.f * s12 + kC2;
        const float B1 = mu1 * mu1 + mu2 * mu2 + kC1, B2 = s11 + s22 + kC2;
        const float inv = 1.f / (B1 * B2);
        const float S = A1 * A2 * inv;
        ss_acc += S;
        l1_acc += fabsf(sx[r + kHalf][c + kHalf] - sy[r + kHalf][c + kHalf]);
        if (dmaps) {
            const float dS_ds11 = -S / B2;                         // d/d sigma1^2  (== d/d E[x^2])
            const float dS_ds12 = 2.f * A1 * inv;                  // d/d sigma12   (== d/d E[xy])
            const float dS_dmu1 = 2.f * mu2 * A2 * inv - 2.f * mu1 * S / B1 - 2.f * mu1 * dS_ds11 - mu2 * dS_ds12;
            const size_t n = (size_t)planes * H * W, o = base + (size_t)gy * W + gx;
            dmaps[o] = dS_dmu1; dmaps[n + o] = dS_ds11; dmaps[2 * n + o] = dS_ds12;
        }
    }
    l1_acc = warp_sum(l1_acc); ss_acc = warp_sum(ss_acc);
    if ((tid & 31) == 0) { red[0][tid >> 5] = l1_acc; red[1][tid >> 5] = ss_acc; }
    __syncthreads();
    const unsigned n_cta = gridDim.x * gridDim.y * gridDim.z;
    const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (tid == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
        partial[cta] = a; partial[n_cta + cta] = b;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == n_cta - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double a = 0, b = 0;
    for (unsigned i = tid; i < n_cta; i += 256) { a += partial[i]; b += partial[n_cta + i]; }   // fixed order per thread
    __shared__ double fin[2][256];
    fin[0][tid] = a; fin[1][tid] = b;
    __syncthreads();
    if (tid == 0) {
        a = 0; b = 0;
        for (int i = 0; i < 256; ++i) { a += fin[0][i]; b += fin[1][i]; }
        const double n = (double)planes * H * W;
        const float l1 = (float)(a / n), ss = (float)(b / n);
        out3[0] = w_l1 * l1 + w_ssim * (1.f - ss);
        out3[1] = l1;
        out3[2] = ss;
        *ticket = 0;                                            // re-arm for the next call on this workspace
    }
}

// dimg[p] = g * ( w_l1 * sign(x-y)/N  -  w_ssim/N * sum_q win(q-p) * (dmu1[q] + 2 x[p] ds11[q] + y[p] ds12[q]) )
__global__ void __launch_bounds__(256) l1_ssim_backward_kernel(int planes, int H, int W, const float* __restrict__ img,
                                                               const float* __restrict__ gt, const float* __restrict__ dmaps,
                                                               const __grid_constant__ Window win, float w_l1, float w_ssim,
                                                               const float* __restrict__ grad_loss, float* __restrict__ dimg) {
    __shared__ float sm[3][kHalo][kHalo + 1];
    __shared__ float hs[3][kHalo][kTile];
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t base = (size_t)plane * H * W, n = (size_t)planes * H * W;
    const int tid = threadIdx.x;
    for (int e = tid; e < kHalo * kHalo; e += 256) {
        const int r = e / kHalo, c = e % kHalo;
        const int gy = y0 + r - kHalf, gx = x0 + c - kHalf;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const size_t o = base + (size_t)gy * W + gx;
#pragma unroll
        for (int q = 0; q < 3; ++q) sm[q][r][c] = in ? dmaps[q * n + o] : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < kHalo * kTile; e += 256) {
        const int r = e / kTile, c = e % kTile;
        float a[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < kWin; ++k)
#pragma unroll
            for (int q = 0; q < 3; ++q) a[q] = fmaf(win.w[k], sm[q][r][c + k], a[q]);
#pragma unroll
        for (int q = 0; q < 3; ++q) hs[q][r][c] = a[q];
    }
    __syncthreads();
    const float g = grad_loss ? *grad_loss : 1.f;
    const float inv_n = (float)(1.0 / ((double)planes * H * W));
    const int c = tid & 31;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (tid >> 5) + 8 * i;
        const int gy = y0 + r, gx = x0 + c;
        if (gy >= H || gx >= W) continue;
        float m[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < kWin; ++k)
#pragma unroll
            for (int q = 0; q < 3; ++q) m[q] = fmaf(win.w[k], hs[q][r + k][c], m[q]);
        const size_t o = base + (size_t)gy * W + gx;
        const float x = img[o], y = gt[o];
        const float d = x - y;
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        const float dss = m[0] + 2.f * x * m[1] + y * m[2];
        dimg[o] = g * inv_n * (w_l1 * sgn - w_ssim * dss);
    }
}

size_t l1_ssim_workspace_bytes(int planes, int H, int W) {
    const size_t n_cta = (size_t)((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile) * (size_t)planes;
    return 2 * n_cta * sizeof(double) + 256;
}

int launch_l1_ssim_fwd(int planes, int H, int W, const float* img, const float* gt, float w_l1, float w_ssim, float* out3,
                       float* dmaps, void* workspace, cudaStream_t stream) {
    static const Window win = make_window();
    dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, planes);
    GPSG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "image too large for the loss grid");
    const size_t n_cta = (size_t)grid.x * grid.y * grid.z;
    double* partial = (double*)workspace;
    unsigned* ticket = (unsigned*)((char*)workspace + 2 * n_cta * sizeof(double));
    GPSG_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), stream));
    l1_ssim_forward_kernel<<<grid, 256, 0, stream>>>(planes, H, W, img, gt, win, w_l1, w_ssim, out3, dmaps, partial, ticket);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

int launch_l1_ssim_bwd(int planes, int H, int W, const float* img, const float* gt, const float* dmaps, float w_l1,
                       float w_ssim, const float* grad_loss, float* dimg, cudaStream_t stream) {
    static const Window win = make_window();
    dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, planes);
    GPSG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "image too large for the loss grid");
    l1_ssim_backward_kernel<<<grid, 256, 0, stream>>>(planes, H, W, img, gt, dmaps, win, w_l1, w_ssim, grad_loss, dimg);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

}  // namespace gpsg
