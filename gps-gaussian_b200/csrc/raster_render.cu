// raster_render.cu -- per-tile front-to-back alpha compositing (upstream
// forward.cu::renderCUDA, behind reference gaussian_renderer/__init__.py:54-62).
//
// B200 design (not upstream's):
//  * a tile's sorted Gaussians are three contiguous float4 slabs (raster_binning.cu) streamed into a shared
//    memory ring by a producer lane with 1-D TMA bulk copies (slab_ring.cuh); consumer warps never block on
//    each other (no per-batch __syncthreads);
//  * a CTA is half a tile (16x8 pixels, 4 consumer warps + 1 producer warp) -> 2x more, smaller work units for
//    the 148 SMs; a warp covers an 8x4 pixel block;
//  * warp-cooperative culling: for every 32 staged Gaussians each lane tests ONE Gaussian's conservative
//    alpha>=1/255 bounding box against the warp's block; a ballot gives the survivors and only those are
//    evaluated by the 32 pixels.  Skipped entries are exactly entries every lane would `continue` on, so the
//    result is unchanged (n_contrib counts list positions, not evaluations);
//  * the conic arrives pre-scaled into the log2 domain, so alpha = o * ex2(p) with p a 5-op polynomial.
#include "gpsg_internal.cuh"
#include "slab_ring.cuh"

namespace gpsg {

constexpr int kFwdChunk = 64;   // Gaussians per ring stage (3 x 1 KB)
constexpr int kFwdStages = 8;
constexpr int kFwdWarps = 4;    // consumer warps per CTA: 16 x 8 pixels

__global__ void __launch_bounds__((kFwdWarps + 1) * 32) render_forward_kernel(const __grid_constant__ Camera cam,
                                                                            const float4* __restrict__ slabA,
                                                                            const float4* __restrict__ slabB,
                                                                            const float4* __restrict__ slabC,
                                                                            const uint2* __restrict__ ranges, const uint32_t* __restrict__ status,
                                                                            float* __restrict__ final_T,
                                                                            uint32_t* __restrict__ n_contrib,
                                                                            float* __restrict__ out_color) {
    __shared__ SlabRing<kFwdChunk, kFwdStages> ring;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_y = blockIdx.y >> 1, half = blockIdx.y & 1;   // two CTAs per 16x16 tile
    const int tile = tile_y * cam.grid_x + blockIdx.x;
    const uint2 range = status[2] ? make_uint2(0u, 0u) : ranges[tile];   // planned-mode overflow: render nothing
    const int total = (int)(range.y - range.x);
    const int nbatch = (total + kFwdChunk - 1) / kFwdChunk;

    if (tid == 0) ring_init(ring, kFwdWarps);
    __syncthreads();

    if (warp == kFwdWarps) {  // ---------------- producer warp ----------------
        if (lane == 0)
            ring_produce(ring, nbatch, kFwdWarps, slabA, slabB, slabC,
                         [&](int b) { return (size_t)range.x + (size_t)b * kFwdChunk; },
                         [&](int b) { return min(kFwdChunk, total - b * kFwdChunk); });
        return;
    }
    // ---------------- consumer warps: warp w covers the 8x4 block at ((w&1)*8, half*8 + (w>>1)*4) ----------------
    const int bx0 = blockIdx.x * GPSG_TILE_X + ((warp & 1) << 3);
    const int by0 = tile_y * GPSG_TILE_Y + (half << 3) + ((warp >> 1) << 2);
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const bool inside = px < cam.W && py < cam.H;
    const float pixfx = (float)px, pixfy = (float)py;
    const float wx0 = (float)bx0, wx1 = (float)(bx0 + 7), wy0 = (float)by0, wy1 = (float)(by0 + 3);

    bool done = !inside;
    bool warp_done = __all_sync(0xffffffffu, done);
    if (warp_done && lane == 0) atomicAdd(&ring.done_warps, 1);
    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
    int last_contributor = 0;

    for (int b = 0; b < nbatch; ++b) {
        if (!ring_wait_full(ring, b, kFwdWarps)) break;
        if (!warp_done) {
            const int s = b % kFwdStages;
            const int n = min(kFwdChunk, total - b * kFwdChunk);
            const float4* __restrict__ SA = ring.A[s];
            const float4* __restrict__ SB = ring.B[s];
            const float4* __restrict__ SC = ring.C[s];
            const int posbase = b * kFwdChunk + 1;
            for (int base = 0; base < n; base += 32) {
                const int my = base + lane;
                bool hit = false;
                if (my < n) {
                    const float4 a = SA[my];
                    hit = (a.x >= wx0 - a.z) && (a.x <= wx1 + a.z) && (a.y >= wy0 - a.w) && (a.y <= wy1 + a.w);
                }
                unsigned m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int j = base + __ffs(m) - 1;
                    m &= m - 1;
                    // Straight-line, predicated evaluation: under SIMT the "contributing" tail runs whenever ANY lane
                    // contributes (almost always for a survivor), so per-lane branches only add BSSY/BSYNC/BRA overhead.
                    const float2 xy = *reinterpret_cast<const float2*>(&SA[j]);
                    const float4 q = SB[j];
                    const float4 c = SC[j];
                    const float dx = xy.x - pixfx, dy = xy.y - pixfy;
                    // p = log2e * power,  power = -0.5*(cx dx^2 + cz dy^2) - cy dx dy
                    const float p = fmaf(q.z * dy, dy, fmaf(q.x, dx, q.y * dy) * dx);
                    const float alpha = fminf(0.99f, q.w * ex2_approx(p));
                    const bool valid = !done && !(p > 0.0f) && !(alpha < 1.0f / 255.0f);
                    const float test_T = T * (1.0f - alpha);
                    const bool stop = valid && (test_T < 0.0001f);
                    const bool upd = valid && !stop;
                    done = done || stop;
                    const float w = upd ? alpha * T : 0.0f;
                    C0 = fmaf(c.x, w, C0);
                    C1 = fmaf(c.y, w, C1);
                    C2 = fmaf(c.z, w, C2);
                    T = upd ? test_T : T;
                    last_contributor = upd ? posbase + j : last_contributor;
                }
                if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }
            }
            if (warp_done && lane == 0) atomicAdd(&ring.done_warps, 1);
        }
        ring_release(ring, b, lane);
    }

    if (inside) {
        const size_t HW = (size_t)cam.W * cam.H;
        const size_t pid = (size_t)py * cam.W + px;
        final_T[pid] = T;
        n_contrib[pid] = (uint32_t)last_contributor;
        out_color[pid] = fmaf(T, cam.bg[0], C0);
        out_color[HW + pid] = fmaf(T, cam.bg[1], C1);
        out_color[2 * HW + pid] = fmaf(T, cam.bg[2], C2);
    }
}

int launch_render_forward(const Camera& cam, BinningState b, ImageState im, float* out_color, cudaStream_t stream) {
    dim3 grid(cam.grid_x, cam.grid_y * 2);
    render_forward_kernel<<<grid, (kFwdWarps + 1) * 32, 0, stream>>>(cam, b.slabA, b.slabB, b.slabC, im.ranges, im.totals,
                                                                    im.final_T, im.n_contrib, out_color);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

}  // namespace gpsg
