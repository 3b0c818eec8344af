// raster_render.cu -- per-tile front-to-back alpha compositing (upstream
// forward.cu::renderCUDA, behind reference gaussian_renderer/__init__.py:54-62).
//
// One CTA per 16x16 tile, one thread per pixel.  The tile's sorted Gaussians are a contiguous
// slab (see raster_binning.cu); a warp covers an 8x4 pixel block (not a 16x2 strip) so that
// whole-warp rejects are frequent.
#include "gpsg_internal.cuh"

namespace gpsg {

// thread -> pixel inside the tile: warp w covers the 8x4 block at ((w&1)*8, (w>>1)*4)
__device__ __forceinline__ void tile_pixel(int tid, int& lx, int& ly) {
    const int w = tid >> 5, l = tid & 31;
    lx = ((w & 1) << 3) + (l & 7);
    ly = ((w >> 1) << 2) + (l >> 3);
}

__global__ void __launch_bounds__(256) render_forward_kernel(const __grid_constant__ Camera cam,
                                                             const float4* __restrict__ slabA,
                                                             const float4* __restrict__ slabB,
                                                             const float4* __restrict__ slabC,
                                                             const uint2* __restrict__ ranges,
                                                             float* __restrict__ final_T,
                                                             uint32_t* __restrict__ n_contrib,
                                                             float* __restrict__ out_color) {
    __shared__ float4 sA[GPSG_TILE_PIX];
    __shared__ float4 sB[GPSG_TILE_PIX];
    __shared__ float sCb[GPSG_TILE_PIX];

    const int tile = blockIdx.y * cam.grid_x + blockIdx.x;
    int lx, ly;
    tile_pixel(threadIdx.x, lx, ly);
    const int px = blockIdx.x * GPSG_TILE_X + lx, py = blockIdx.y * GPSG_TILE_Y + ly;
    const bool inside = px < cam.W && py < cam.H;
    const float pixfx = (float)px, pixfy = (float)py;
    const uint2 range = ranges[tile];
    int todo = (int)(range.y - range.x);
    const int rounds = (todo + GPSG_TILE_PIX - 1) / GPSG_TILE_PIX;

    bool done = !inside;
    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
    uint32_t contributor = 0, last_contributor = 0;

    for (int r = 0; r < rounds; ++r, todo -= GPSG_TILE_PIX) {
        if (__syncthreads_count(done) == GPSG_TILE_PIX) break;
        const int n = min(GPSG_TILE_PIX, todo);
        if ((int)threadIdx.x < n) {
            const size_t k = (size_t)range.x + (size_t)r * GPSG_TILE_PIX + threadIdx.x;
            sA[threadIdx.x] = slabA[k];
            sB[threadIdx.x] = slabB[k];
            sCb[threadIdx.x] = slabC[k].x;
        }
        __syncthreads();
        for (int j = 0; !done && j < n; ++j) {
            ++contributor;
            const float4 a = sA[j];
            const float4 b = sB[j];
            const float dx = a.x - pixfx, dy = a.y - pixfy;
            const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, b.y * __expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const float w = alpha * T;
            C0 += b.z * w;
            C1 += b.w * w;
            C2 += sCb[j] * w;
            T = test_T;
            last_contributor = contributor;
        }
    }
    if (inside) {
        const size_t HW = (size_t)cam.W * cam.H;
        const size_t pid = (size_t)py * cam.W + px;
        final_T[pid] = T;
        n_contrib[pid] = last_contributor;
        out_color[pid] = C0 + T * cam.bg[0];
        out_color[HW + pid] = C1 + T * cam.bg[1];
        out_color[2 * HW + pid] = C2 + T * cam.bg[2];
    }
}

int launch_render_forward(const Camera& cam, BinningState b, ImageState im, float* out_color, cudaStream_t stream) {
    dim3 grid(cam.grid_x, cam.grid_y);
    render_forward_kernel<<<grid, GPSG_TILE_PIX, 0, stream>>>(cam, b.slabA, b.slabB, b.slabC, im.ranges, im.final_T,
                                                             im.n_contrib, out_color);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

}  // namespace gpsg
