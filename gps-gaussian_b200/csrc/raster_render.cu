// raster_render.cu -- per-tile front-to-back alpha compositing (upstream
// forward.cu::renderCUDA, behind reference gaussian_renderer/__init__.py:54-62).
//
// B200 design (not upstream's):
//  * a tile's sorted Gaussians are three contiguous float4 slabs (raster_binning.cu); each 256-entry batch is
//    brought into shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), double-buffered,
//    so the next batch streams in while the current one is composited;
//  * a warp covers an 8x4 pixel block.  For every 32 staged Gaussians each lane tests ONE Gaussian's conservative
//    alpha>=1/255 bounding box against the warp's block; a ballot gives the survivors and only those are evaluated
//    by all 32 pixels.  Skipped entries are exactly entries every lane would `continue` on, so results are
//    unchanged (n_contrib counts list positions, not evaluations).
#include <cstdlib>
#include "gpsg_internal.cuh"
#include "tma_bulk.cuh"

namespace gpsg {

constexpr int kBatch = GPSG_TILE_PIX;  // Gaussians per shared-memory stage
constexpr int kStages = 2;

struct __align__(128) RenderStage {
    float4 A[kBatch];  // x, y, cull half-extent x, y
    float4 B[kBatch];  // conic.x, conic.y, conic.z, opacity
    float4 C[kBatch];  // r, g, b, id
};

__global__ void __launch_bounds__(256) render_forward_kernel(const __grid_constant__ Camera cam,
                                                             const float4* __restrict__ slabA,
                                                             const float4* __restrict__ slabB,
                                                             const float4* __restrict__ slabC,
                                                             const uint2* __restrict__ ranges,
                                                             float* __restrict__ final_T,
                                                             uint32_t* __restrict__ n_contrib,
                                                             float* __restrict__ out_color) {
    __shared__ RenderStage st[kStages];
    __shared__ __align__(8) uint64_t full_bar[kStages];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * cam.grid_x + blockIdx.x;
    // warp w covers the 8x4 pixel block at ((w&1)*8, (w>>1)*4) of the tile
    const int bx0 = blockIdx.x * GPSG_TILE_X + ((warp & 1) << 3), by0 = blockIdx.y * GPSG_TILE_Y + ((warp >> 1) << 2);
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const bool inside = px < cam.W && py < cam.H;
    const float pixfx = (float)px, pixfy = (float)py;
    const float wx0 = (float)bx0, wx1 = (float)(bx0 + 7), wy0 = (float)by0, wy1 = (float)(by0 + 3);
    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);
    const int nbatch = (total + kBatch - 1) / kBatch;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int b) {  // thread 0 only
        const int s = b % kStages;
        const int n = min(kBatch, total - b * kBatch);
        const uint32_t bytes = (uint32_t)n * 16u;
        const size_t k = (size_t)range.x + (size_t)b * kBatch;
        mbar_expect_tx(&full_bar[s], 3u * bytes);
        tma_bulk_g2s(st[s].A, slabA + k, bytes, &full_bar[s]);
        tma_bulk_g2s(st[s].B, slabB + k, bytes, &full_bar[s]);
        tma_bulk_g2s(st[s].C, slabC + k, bytes, &full_bar[s]);
    };
    if (tid == 0) {
        if (nbatch > 0) issue(0);
        if (nbatch > 1) issue(1);
    }

    bool done = !inside;
    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
    uint32_t last_contributor = 0;
    int b = 0;
    for (; b < nbatch; ++b) {
        const int s = b % kStages;
        mbar_wait(&full_bar[s], (uint32_t)((b / kStages) & 1));
        const int n = min(kBatch, total - b * kBatch);
        const RenderStage& S = st[s];
        for (int base = 0; base < n; base += 32) {
            if (__all_sync(0xffffffffu, done)) break;
            const int my = base + lane;
            bool hit = false;
            if (my < n) {
                const float4 a = S.A[my];
                hit = (a.x >= wx0 - a.z) && (a.x <= wx1 + a.z) && (a.y >= wy0 - a.w) && (a.y <= wy1 + a.w);
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int j = base + __ffs(m) - 1;
                m &= m - 1;
                if (done) continue;
                const float4 a = S.A[j];
                const float4 q = S.B[j];
                const float dx = a.x - pixfx, dy = a.y - pixfy;
                const float power = -0.5f * (q.x * dx * dx + q.z * dy * dy) - q.y * dx * dy;
                if (power > 0.0f) continue;
                const float alpha = fminf(0.99f, q.w * __expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                const float test_T = T * (1.0f - alpha);
                if (test_T < 0.0001f) { done = true; continue; }
                const float4 c = S.C[j];
                const float w = alpha * T;
                C0 += c.x * w;
                C1 += c.y * w;
                C2 += c.z * w;
                T = test_T;
                last_contributor = (uint32_t)(b * kBatch + j + 1);
            }
        }
        // everyone has finished reading stage s -> it may be refilled; also the block-wide early-out vote
        if (__syncthreads_count(done) == GPSG_TILE_PIX) break;
        if (tid == 0 && b + kStages < nbatch) issue(b + kStages);
    }
    // early exit with a bulk copy still in flight into this CTA's shared memory: wait for it before leaving
    if (tid == 0 && b < nbatch && b + 1 < nbatch) mbar_wait(&full_bar[(b + 1) % kStages], (uint32_t)(((b + 1) / kStages) & 1));

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

// ---- v1 (kept for A/B timing only: GPSG_RENDER_IMPL=v1): cooperative loads, no culling, no TMA ----------------
__global__ void __launch_bounds__(256) render_forward_v1_kernel(const __grid_constant__ Camera cam,
                                                                const float4* __restrict__ slabA,
                                                                const float4* __restrict__ slabB,
                                                                const float4* __restrict__ slabC,
                                                                const uint2* __restrict__ ranges,
                                                                float* __restrict__ final_T,
                                                                uint32_t* __restrict__ n_contrib,
                                                                float* __restrict__ out_color) {
    __shared__ float4 sA[GPSG_TILE_PIX];
    __shared__ float4 sB[GPSG_TILE_PIX];
    __shared__ float4 sC[GPSG_TILE_PIX];
    const int tile = blockIdx.y * cam.grid_x + blockIdx.x;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int px = blockIdx.x * GPSG_TILE_X + ((w & 1) << 3) + (l & 7), py = blockIdx.y * GPSG_TILE_Y + ((w >> 1) << 2) + (l >> 3);
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
            sC[threadIdx.x] = slabC[k];
        }
        __syncthreads();
        for (int j = 0; !done && j < n; ++j) {
            ++contributor;
            const float4 a = sA[j];
            const float4 q = sB[j];
            const float dx = a.x - pixfx, dy = a.y - pixfy;
            const float power = -0.5f * (q.x * dx * dx + q.z * dy * dy) - q.y * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, q.w * __expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const float4 c = sC[j];
            const float wgt = alpha * T;
            C0 += c.x * wgt; C1 += c.y * wgt; C2 += c.z * wgt;
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

bool use_v1_kernels() {
    static const bool v1 = [] { const char* e = getenv("GPSG_RENDER_IMPL"); return e && e[0] == 'v' && e[1] == '1'; }();
    return v1;
}

int launch_render_forward(const Camera& cam, BinningState b, ImageState im, float* out_color, cudaStream_t stream) {
    dim3 grid(cam.grid_x, cam.grid_y);
    if (use_v1_kernels())
        render_forward_v1_kernel<<<grid, GPSG_TILE_PIX, 0, stream>>>(cam, b.slabA, b.slabB, b.slabC, im.ranges,
                                                                    im.final_T, im.n_contrib, out_color);
    else
        render_forward_kernel<<<grid, GPSG_TILE_PIX, 0, stream>>>(cam, b.slabA, b.slabB, b.slabC, im.ranges, im.final_T,
                                                                 im.n_contrib, out_color);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

}  // namespace gpsg
