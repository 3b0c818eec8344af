// raster_binning.cu -- tile binning (upstream rasterizer_impl.cu:
// InclusiveSum -> duplicateWithKeys -> SortPairs -> identifyTileRanges).
//
// B200 redesign of the last step: after the sort, one pass gathers every pair's render parameters
// into three tile-contiguous float4 "slab" arrays, so the compositing kernels stream each tile's
// list with 1-D TMA bulk copies instead of chasing point_list[] indirections, and the backward
// re-uses the same slabs.  The same pass detects tile boundaries (identifyTileRanges).
#include <cub/cub.cuh>
#include "gpsg_internal.cuh"

namespace gpsg {

size_t scan_temp_bytes(int P) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, P > 0 ? P : 1);
    return bytes;
}

size_t sort_temp_bytes(size_t N, int end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int64_t)(N > 0 ? N : 1), 0, end_bit);
    return bytes;
}

int run_scan(GeomState g, int P, cudaStream_t stream) {
    if (P <= 0) return GPSG_OK;
    size_t bytes = g.scan_temp_bytes;
    GPSG_CUDA(cub::DeviceScan::InclusiveSum(g.scan_temp, bytes, g.tiles_touched, g.point_offsets, P, stream));
    return GPSG_OK;
}

// One thread per Gaussian; emits its (tile<<32 | depth bits, id) pairs in row-major rect order.
__global__ void __launch_bounds__(256) duplicate_kernel(const __grid_constant__ Camera cam, int P,
                                                        const int32_t* __restrict__ radii, GeomState g,
                                                        BinningState b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int radius = radii[i];
    if (radius <= 0) return;
    uint32_t off = (i == 0) ? 0u : g.point_offsets[i - 1];
    const float2 p = g.means2D[i];
    const float rad = (float)radius;
    // identical expressions to preprocess (pure scaling by 1/16 and casts: exact)
    const int rx0 = min(cam.grid_x, max(0, (int)((p.x - rad) / (float)GPSG_TILE_X)));
    const int ry0 = min(cam.grid_y, max(0, (int)((p.y - rad) / (float)GPSG_TILE_Y)));
    const int rx1 = min(cam.grid_x, max(0, (int)((p.x + rad + (float)(GPSG_TILE_X - 1)) / (float)GPSG_TILE_X)));
    const int ry1 = min(cam.grid_y, max(0, (int)((p.y + rad + (float)(GPSG_TILE_Y - 1)) / (float)GPSG_TILE_Y)));
    const uint64_t dbits = (uint64_t)__float_as_uint(g.depths[i]);
    for (int y = ry0; y < ry1; ++y)
        for (int x = rx0; x < rx1; ++x) {
            const uint64_t key = ((uint64_t)(uint32_t)(y * cam.grid_x + x) << 32) | dbits;
            b.keys_unsorted[off] = key;
            b.vals_unsorted[off] = (uint32_t)i;
            ++off;
        }
}

int launch_duplicate(const Camera& cam, int P, const int32_t* radii, GeomState g, BinningState b, cudaStream_t stream) {
    if (P <= 0) return GPSG_OK;
    duplicate_kernel<<<(P + 255) / 256, 256, 0, stream>>>(cam, P, radii, g, b);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

int run_sort(BinningState b, size_t N, int end_bit, cudaStream_t stream) {
    if (N == 0) return GPSG_OK;
    size_t bytes = b.sort_temp_bytes;
    GPSG_CUDA(cub::DeviceRadixSort::SortPairs(b.sort_temp, bytes, b.keys_unsorted, b.keys, b.vals_unsorted, b.vals,
                                              (int64_t)N, 0, end_bit, stream));
    return GPSG_OK;
}

// One thread per sorted pair: tile-range detection + parameter gather into the slabs.
__global__ void __launch_bounds__(256) gather_ranges_kernel(size_t N, const float* __restrict__ colors, GeomState g,
                                                            BinningState b, ImageState im) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint64_t key = b.keys[i];
    const uint32_t tile = (uint32_t)(key >> 32);
    if (i == 0) {
        im.ranges[tile].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(b.keys[i - 1] >> 32);
        if (prev != tile) {
            im.ranges[prev].y = (uint32_t)i;
            im.ranges[tile].x = (uint32_t)i;
        }
    }
    if (i == N - 1) im.ranges[tile].y = (uint32_t)N;
    const uint32_t id = b.vals[i];
    const float2 xy = g.means2D[id];
    const float4 co = g.conic_opacity[id];
    const float r = colors[3 * id], gg = colors[3 * id + 1], bb = colors[3 * id + 2];
    // Conservative screen-space half-extents of the region where alpha = o*exp(power) can reach 1/255:
    // power >= -tau, tau = ln(255 o)  <=>  d^T Conic d <= 2 tau  -> bounding box sqrt(2 tau * Sigma_xx/yy),
    // Sigma = Conic^-1.  Used only to SKIP work that the per-pixel tests would reject anyway (results unchanged).
    float ex = 3.0e38f, ey = 3.0e38f;                       // degenerate conic: never cull
    const float detc = co.x * co.z - co.y * co.y;
    if (!(co.w * 255.0f >= 1.0f)) {
        ex = ey = -3.0e38f;                                  // alpha < 1/255 everywhere: always culled
    } else if (detc > 0.0f && co.x > 0.0f && co.z > 0.0f) {
        const float tau2 = 2.0f * __logf(co.w * 255.0f) * 1.0005f + 1e-4f;
        ex = sqrtf(tau2 * co.z / detc) * 1.0005f + 0.01f;
        ey = sqrtf(tau2 * co.x / detc) * 1.0005f + 0.01f;
        if (!(ex == ex) || !(ey == ey)) { ex = 3.0e38f; ey = 3.0e38f; }
    }
    b.slabA[i] = make_float4(xy.x, xy.y, ex, ey);
    // conic pre-scaled into the log2 domain: alpha = o * 2^(Bx dx^2 + By dx dy + Bz dy^2)
    const float kL = 1.4426950408889634f;
    b.slabB[i] = make_float4(-0.5f * kL * co.x, -kL * co.y, -0.5f * kL * co.z, co.w);
    b.slabC[i] = make_float4(r, gg, bb, __uint_as_float(id));
}

int launch_gather_ranges(const Camera& cam, size_t N, const float* colors, GeomState g, BinningState b, ImageState im,
                         cudaStream_t stream) {
    GPSG_CUDA(cudaMemsetAsync(im.ranges, 0, sizeof(uint2) * (size_t)cam.grid_x * cam.grid_y, stream));
    if (N == 0) return GPSG_OK;
    gather_ranges_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(N, colors, g, b, im);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

}  // namespace gpsg
