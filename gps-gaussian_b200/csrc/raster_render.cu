// raster_render.cu -- per-tile front-to-back alpha compositing (upstream
// forward.cu::renderCUDA, behind reference gaussian_renderer/__init__.py:54-62).
//
// B200 design (not upstream's):
//  * a tile's sorted Gaussians are three contiguous float4 slabs (raster_binning.cu) streamed into a shared
//    memory ring by a producer lane with 1-D TMA bulk copies (slab_ring.cuh); consumer warps never block on
//    each other (no per-batch __syncthreads);
//  * a CTA is one 16x16 tile (4 consumer warps + 1 producer warp); a warp covers an 8x8 pixel block with TWO pixels
//    per lane, evaluated with packed fp32 (FFMA2/FMUL2/FADD2): half the issue slots per pixel;
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
constexpr int kFwdWarps = 4;    // consumer warps per CTA: one 16 x 16 tile, each warp an 8 x 8 block, 2 pixels per lane

// Two pixels per lane -- (x, y) and (x, y+4) of the warp's 8x8 block -- so that the per-pixel arithmetic runs on
// packed fp32 (FFMA2/FMUL2/FADD2): the kernel is issue-slot-bound and a packed instruction does two pixels' work in
// one slot.  Component .x of every float2 below belongs to the upper pixel, .y to the lower one.
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
    const int tile = blockIdx.y * cam.grid_x + blockIdx.x;
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
    // ---------------- consumer warps: warp w covers the 8x8 block at ((w&1)*8, (w>>1)*8) of the tile ----------------
    const int bx0 = blockIdx.x * GPSG_TILE_X + ((warp & 1) << 3);
    const int by0 = blockIdx.y * GPSG_TILE_Y + ((warp >> 1) << 3);
    const int px = bx0 + (lane & 7), py0 = by0 + (lane >> 3), py1 = py0 + 4;
    const bool in0 = px < cam.W && py0 < cam.H, in1 = px < cam.W && py1 < cam.H;
    const float pixfx = (float)px;
    const float2 npy = make_float2(-(float)py0, -(float)py1);
    const float wx0 = (float)bx0, wx1 = (float)(bx0 + 7), wy0 = (float)by0, wy1 = (float)(by0 + 7);

    bool done0 = !in0, done1 = !in1;
    bool warp_done = __all_sync(0xffffffffu, done0 && done1);
    if (warp_done && lane == 0) atomicAdd(&ring.done_warps, 1);
    float2 T = make_float2(1.0f, 1.0f), C0 = make_float2(0.f, 0.f), C1 = C0, C2 = C0;
    int last0 = 0, last1 = 0;

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
                    // Straight-line, predicated evaluation of both pixels (no per-lane branches: under SIMT the tail
                    // runs whenever any lane contributes, so branches would only add BSSY/BSYNC/BRA overhead).
                    const float2 xy = *reinterpret_cast<const float2*>(&SA[j]);
                    const float4 q = SB[j];
                    const float4 c = SC[j];
                    const float dx = xy.x - pixfx;
                    const float2 dy = add2(splat2(xy.y), npy);
                    // p = log2e * power = Bz dy^2 + (Bx dx + By dy) dx      (same operation order as the scalar form)
                    const float2 u = fma2(splat2(q.x), splat2(dx), mul2(splat2(q.y), dy));
                    const float2 p = fma2(mul2(splat2(q.z), dy), dy, mul2(u, splat2(dx)));
                    const float2 og = mul2(splat2(q.w), make_float2(ex2_approx(p.x), ex2_approx(p.y)));
                    const float2 alpha = make_float2(fminf(0.99f, og.x), fminf(0.99f, og.y));
                    const bool v0 = !done0 && !(p.x > 0.0f) && !(alpha.x < 1.0f / 255.0f);
                    const bool v1 = !done1 && !(p.y > 0.0f) && !(alpha.y < 1.0f / 255.0f);
                    const float2 test_T = mul2(T, add2(splat2(1.0f), make_float2(-alpha.x, -alpha.y)));
                    const bool stop0 = v0 && (test_T.x < 0.0001f), stop1 = v1 && (test_T.y < 0.0001f);
                    const bool u0 = v0 && !stop0, u1 = v1 && !stop1;
                    done0 = done0 || stop0;
                    done1 = done1 || stop1;
                    const float2 aT = mul2(alpha, T);
                    const float2 w = make_float2(u0 ? aT.x : 0.0f, u1 ? aT.y : 0.0f);
                    C0 = fma2(splat2(c.x), w, C0);
                    C1 = fma2(splat2(c.y), w, C1);
                    C2 = fma2(splat2(c.z), w, C2);
                    T = make_float2(u0 ? test_T.x : T.x, u1 ? test_T.y : T.y);
                    last0 = u0 ? posbase + j : last0;
                    last1 = u1 ? posbase + j : last1;
                }
                if (__all_sync(0xffffffffu, done0 && done1)) { warp_done = true; break; }
            }
            if (warp_done && lane == 0) atomicAdd(&ring.done_warps, 1);
        }
        ring_release(ring, b, lane);
    }

    const size_t HW = (size_t)cam.W * cam.H;
    if (in0) {
        const size_t pid = (size_t)py0 * cam.W + px;
        final_T[pid] = T.x;
        n_contrib[pid] = (uint32_t)last0;
        out_color[pid] = fmaf(T.x, cam.bg[0], C0.x);
        out_color[HW + pid] = fmaf(T.x, cam.bg[1], C1.x);
        out_color[2 * HW + pid] = fmaf(T.x, cam.bg[2], C2.x);
    }
    if (in1) {
        const size_t pid = (size_t)py1 * cam.W + px;
        final_T[pid] = T.y;
        n_contrib[pid] = (uint32_t)last1;
        out_color[pid] = fmaf(T.y, cam.bg[0], C0.y);
        out_color[HW + pid] = fmaf(T.y, cam.bg[1], C1.y);
        out_color[2 * HW + pid] = fmaf(T.y, cam.bg[2], C2.y);
    }
}

int launch_render_forward(const Camera& cam, BinningState b, ImageState im, float* out_color, cudaStream_t stream) {
    dim3 grid(cam.grid_x, cam.grid_y);
    render_forward_kernel<<<grid, (kFwdWarps + 1) * 32, 0, stream>>>(cam, b.slabA, b.slabB, b.slabC, im.ranges, im.totals,
                                                                    im.final_T, im.n_contrib, out_color);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

}  // namespace gpsg
