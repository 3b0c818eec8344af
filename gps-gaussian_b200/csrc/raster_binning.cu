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


// =====================================================================================================
// Tile-bucket binning (default path).  Replaces InclusiveSum + duplicateWithKeys + a 6-pass global radix sort +
// identifyTileRanges + gather by:  [preprocess counts pairs per tile] -> tile_scan -> bucket_scatter ->
// tile_sort_gather.  The order inside a tile is defined by the 64-bit key (depth bits << 32 | Gaussian id), which is
// exactly the order a STABLE sort of (tile, depth) keys emitted in Gaussian-index order produces (a Gaussian emits at
// most one pair per tile), so point lists, keys and ranges are bit-identical to the radix path and to the oracle.
// =====================================================================================================

// one CTA: exclusive scan of the per-tile counts -> ranges[t] = (start, end); totals = (N, max count); cursors = 0
__global__ void __launch_bounds__(1024) tile_scan_kernel(int tiles, ImageState im) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry, smax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry = 0; smax = 0; }
    __syncthreads();
    uint32_t local_max = 0;
    for (int base = 0; base < tiles; base += 1024) {
        const int t = base + tid;
        const uint32_t c = t < tiles ? im.tile_count[t] : 0u;
        local_max = max(local_max, c);
        uint32_t v = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        if (lane == 31) warp_sums[warp] = v;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = v + (warp ? warp_sums[warp - 1] : 0u) + carry;
        if (t < tiles) {
            im.ranges[t] = make_uint2(incl - c, incl);
            im.tile_cursor[t] = 0u;
        }
        __syncthreads();
        if (tid == 1023) carry = incl;
        __syncthreads();
    }
    local_max = __reduce_max_sync(0xffffffffu, local_max);
    if (lane == 0) atomicMax(&smax, local_max);
    __syncthreads();
    if (tid == 0) { im.totals[0] = carry; im.totals[1] = smax; }
}

int launch_tile_scan(const Camera& cam, ImageState im, cudaStream_t stream) {
    tile_scan_kernel<<<1, 1024, 0, stream>>>(cam.grid_x * cam.grid_y, im);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

// one thread per Gaussian: claims a slot in each touched tile's bucket and writes (depth bits, id)
__global__ void __launch_bounds__(256) bucket_scatter_kernel(const __grid_constant__ Camera cam, int P,
                                                             const int32_t* __restrict__ radii, GeomState g,
                                                             BinningState b, ImageState im) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int radius = radii[i];
    if (radius <= 0) return;
    const float2 p = g.means2D[i];
    const float rad = (float)radius;
    const int rx0 = min(cam.grid_x, max(0, (int)((p.x - rad) / (float)GPSG_TILE_X)));
    const int ry0 = min(cam.grid_y, max(0, (int)((p.y - rad) / (float)GPSG_TILE_Y)));
    const int rx1 = min(cam.grid_x, max(0, (int)((p.x + rad + (float)(GPSG_TILE_X - 1)) / (float)GPSG_TILE_X)));
    const int ry1 = min(cam.grid_y, max(0, (int)((p.y + rad + (float)(GPSG_TILE_Y - 1)) / (float)GPSG_TILE_Y)));
    const uint2 entry = make_uint2((uint32_t)i, __float_as_uint(g.depths[i]));   // little-endian u64 = depth<<32 | id
    for (int y = ry0; y < ry1; ++y)
        for (int x = rx0; x < rx1; ++x) {
            const int t = y * cam.grid_x + x;
            const uint32_t pos = atomicAdd(&im.tile_cursor[t], 1u);
            b.bucket[im.ranges[t].x + pos] = entry;
        }
}

int launch_bucket_scatter(const Camera& cam, int P, const int32_t* radii, GeomState g, BinningState b, ImageState im,
                          cudaStream_t stream) {
    if (P <= 0) return GPSG_OK;
    bucket_scatter_kernel<<<(P + 255) / 256, 256, 0, stream>>>(cam, P, radii, g, b, im);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

__device__ __forceinline__ void slab_entry(uint32_t id, const float* __restrict__ colors, const GeomState& g, float4& A,
                                           float4& B, float4& C) {
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
    A = make_float4(xy.x, xy.y, ex, ey);
    // conic pre-scaled into the log2 domain: alpha = o * 2^(Bx dx^2 + By dx dy + Bz dy^2)
    const float kL = 1.4426950408889634f;
    B = make_float4(-0.5f * kL * co.x, -kL * co.y, -0.5f * kL * co.z, co.w);
    C = make_float4(r, gg, bb, __uint_as_float(id));
}

// one CTA per tile: bitonic sort of the tile's bucket in shared memory on (depth bits << 32 | id), then the sorted
// point list, keys and parameter slabs are written (coalesced) -- the sort and the gather never round-trip to HBM.
__global__ void __launch_bounds__(256) tile_sort_gather_kernel(const float* __restrict__ colors, GeomState g,
                                                               BinningState b, ImageState im) {
    extern __shared__ __align__(16) unsigned long long skeys[];
    const int tile = blockIdx.x;
    const uint2 range = im.ranges[tile];
    const int n = (int)(range.y - range.x);
    if (n == 0) return;
    int np = 32;
    while (np < n) np <<= 1;
    const unsigned long long* __restrict__ src = reinterpret_cast<const unsigned long long*>(b.bucket) + range.x;
    for (int i = threadIdx.x; i < np; i += blockDim.x) skeys[i] = i < n ? src[i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= np; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int c = threadIdx.x; c < (np >> 1); c += blockDim.x) {
                const int a = ((c & ~(j - 1)) << 1) | (c & (j - 1));
                const int d = a | j;
                const unsigned long long ka = skeys[a], kd = skeys[d];
                const bool up = (a & k) == 0;
                if ((ka > kd) == up) { skeys[a] = kd; skeys[d] = ka; }
            }
            __syncthreads();
        }
    }
    const unsigned long long tile_hi = (unsigned long long)(uint32_t)tile << 32;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long key = skeys[i];
        const uint32_t id = (uint32_t)key;
        const size_t o = (size_t)range.x + i;
        b.keys[o] = tile_hi | (key >> 32);
        b.vals[o] = id;
        float4 A, B, C;
        slab_entry(id, colors, g, A, B, C);
        b.slabA[o] = A;
        b.slabB[o] = B;
        b.slabC[o] = C;
    }
}

int launch_tile_sort_gather(const Camera& cam, uint32_t max_count, const float* colors, GeomState g, BinningState b,
                            ImageState im, cudaStream_t stream) {
    uint32_t np = 32;
    while (np < max_count) np <<= 1;
    const size_t smem = (size_t)np * sizeof(unsigned long long);
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        GPSG_CUDA(cudaFuncSetAttribute(tile_sort_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set = true;
    }
    tile_sort_gather_kernel<<<cam.grid_x * cam.grid_y, 256, smem, stream>>>(colors, g, b, im);
    GPSG_LAUNCH_CHECK();
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
    float4 A, B, C;
    slab_entry(id, colors, g, A, B, C);
    b.slabA[i] = A;
    b.slabB[i] = B;
    b.slabC[i] = C;
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
