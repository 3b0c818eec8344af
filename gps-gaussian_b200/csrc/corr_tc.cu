// corr_tc.cu -- the fp16 correlation-volume build on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same contract as corr_build_kernel<__half> in corr.cu (reference core/corr.py:53-61 under stage-2 AMP: volume =
// einsum('aijk,aijh->ajkh', f1, f2) / sqrt(D), then the avg_pool2d([1,2]) pyramid of core/corr.py:36-42, every level
// rounded to fp16 like the torch op chain).  Per (b, h) the volume is the GEMM  C[x, y] = sum_d F1[d, x] * F2[d, y]:
// M = W1 (tiles of 128), N = W2 (<= 128), K = D.  Both operands are "MN-major" in the UMMA sense -- the feature maps
// are [B, D, H, W] with W contiguous, so for a fixed (b, h) the M (resp. N) index is the contiguous one -- which
// tcgen05.mma.kind::f16 accepts directly (instruction-descriptor a_major = b_major = 1): no transpose anywhere.
//
// One CTA (256 threads) per (M tile, b, h):
//   1. all warps copy the [D x 128] / [D x N] operand panels global -> shared in 16-byte chunks, placing them in the
//      no-swizzle canonical layout (8 x 8 "core matrices" of 128 contiguous bytes: 8 K rows x 16 B of 8 MN elements;
//      MN-adjacent cores 128 B apart (SBO), K-adjacent core groups one panel row of cores apart (LBO));
//   2. ONE thread issues D/16 tcgen05.mma (UMMA 128 x N x 16, fp32 accumulate in TMEM) and a tcgen05.commit on an
//      mbarrier;
//   3. all 8 warps read their TMEM quadrant back with tcgen05.ld (32 lanes x 16 columns per instruction; thread = one
//      volume row x, 16 consecutive y), scale / round / pool down the pyramid in registers and store 16-byte vectors.
// 98 KB of shared memory and 128 TMEM columns per CTA at D = 192: two CTAs per SM, so one CTA's loads overlap the
// other's MMA + epilogue.  HBM-bound by design: 2*D*W*2 B in, 1.875*W1*W2*2 B out per (b, h).
#include "gpsg_internal.cuh"
#include "tma_bulk.cuh"

#include <cuda_fp16.h>

namespace gpsg {

namespace {

constexpr int kTcThreads = 256;
constexpr int kTcM = 128;              // UMMA M (rows of the volume per CTA)
constexpr uint32_t kTmemCols = 128;    // power of two >= N

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, no swizzle (layout_type 0), version 1 (Blackwell); offsets in 16-byte units
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes) : "memory");
}

// correctly rounded a / d from r = RN(1/d) with one Newton correction (Markstein): q0 = a r; q = q0 + (a - q0 d) r.
// Valid here: |a| is an fp16 value (no overflow / underflow in the residual), d = sqrt(D) is a normal fp32 number.
__device__ __forceinline__ float div_rn_fast(float a, float d, float r) {
    const float q0 = a * r;
    return fmaf(fmaf(-q0, d, a), r, q0);
}

__device__ __forceinline__ float rh(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __halves2half2(__float2half_rn(a), __float2half_rn(b));
    return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(kTcThreads) corr_build_tc_kernel(int D, int H, int W1, int W2,
                                                                   const __half* __restrict__ f1,
                                                                   const __half* __restrict__ f2, __half* __restrict__ v0,
                                                                   __half* __restrict__ v1, __half* __restrict__ v2,
                                                                   __half* __restrict__ v3, int levels, float div) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x_base = blockIdx.x * kTcM;
    const int b = blockIdx.y / H, h = blockIdx.y % H;
    const int NB = W2 >> 3;                                     // N cores per K group
    const int KC = D >> 3;                                      // K core groups
    uint8_t* sA = smem;                                         // [KC][16 cores][128 B]
    uint8_t* sB = smem + (size_t)KC * 2048;                     // [KC][NB cores][128 B]

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }

    // ---- operand panels -> shared, canonical no-swizzle MN-major layout --------------------------------------------
    // a warp moves one K core group (8 d rows) x 4 chunks of 8 elements: lanes 0-7 -> rows of chunk 0 (128 contiguous
    // bytes of shared memory, conflict-free), lanes 8-15 chunk 1, ...; per row the 4 chunks are 64 contiguous bytes of
    // global memory.
    const size_t plane1 = (size_t)H * W1, plane2 = (size_t)H * W2;
    const __half* g1 = f1 + (size_t)b * D * plane1 + (size_t)h * W1 + x_base;
    const __half* g2 = f2 + (size_t)b * D * plane2 + (size_t)h * W2;
    const int cl = lane >> 3, dl = lane & 7;
    // cp.async (LDGSTS): every thread puts all of its 16-byte copies in flight before waiting on any of them -- the whole
    // 2*D*W*2-byte panel pair is outstanding at once (register-staged loads stalled on the scoreboard after 4).
    for (int it = warp; it < KC * 4; it += kTcThreads / 32) {
        const int kc = it >> 2, c = (it & 3) * 4 + cl;
        const bool in = x_base + c * 8 < W1;                      // ragged last M tile: zero-fill (src-size 0)
        cp_async16(sA + (size_t)kc * 2048 + c * 128 + dl * 16, in ? g1 + (size_t)(kc * 8 + dl) * plane1 + c * 8 : g1, in ? 16u : 0u);
    }
    const int NBg = (NB + 3) >> 2;
    for (int it = warp; it < KC * NBg; it += kTcThreads / 32) {
        const int kc = it / NBg, c = (it % NBg) * 4 + cl;
        if (c < NB) cp_async16(sB + ((size_t)kc * NB + c) * 128 + dl * 16, g2 + (size_t)(kc * 8 + dl) * plane2 + c * 8, 16u);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    // ---- MMA: one thread, D/16 instructions of 128 x N x 16 ------------------------------------------------------------
    if (tid == 0) {
        // instruction descriptor: D fp32 (bits 4-5 = 1), A/B fp16 (0), A and B MN-major (bits 15, 16), N>>3 at 17, M>>4 at 24
        const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(W2 >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        const uint32_t lbo_b = (uint32_t)NB * 128u;
        for (int s = 0; s < (D >> 4); ++s) {
            const uint64_t da = umma_desc(a0 + (uint32_t)s * 2u * 2048u, 2048u, 128u);
            const uint64_t db = umma_desc(b0 + (uint32_t)s * 2u * lbo_b, lbo_b, 128u);
            umma_f16(tmem, da, db, idesc, s > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                     : "memory");
    }
    mbar_wait(&bar, 0);
    tc_fence_after();

    // ---- epilogue: TMEM -> registers -> scale / round / pool -> global -------------------------------------------------
    const int quad = warp & 3;                                  // TMEM lanes 32*quad .. +31 belong to this warp
    const int x = x_base + quad * 32 + lane;
    const size_t row = ((size_t)b * H + h) * W1 + x;
    const int Wl1 = W2 >> 1, Wl2 = W2 >> 2, Wl3 = W2 >> 3;
    const float rdiv = __frcp_rn(div);
    for (int j = warp >> 2; j < (W2 >> 4); j += 2) {            // 16-column batches, split between the two warp sets
        float acc[16];
        __syncwarp();                                            // tcgen05.ld is .sync.aligned: whole warp, converged
        tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * 16), acc);
        if (x < W1) {
        float q0[16], q1[8], q2[4], q3[2];
#pragma unroll
        for (int i = 0; i < 16; ++i) q0[i] = rh(div_rn_fast(rh(acc[i]), div, rdiv));   // einsum result in fp16, then the division in fp16
        uint4* o0 = reinterpret_cast<uint4*>(v0 + row * W2 + j * 16);
        o0[0] = make_uint4(pack2(q0[0], q0[1]), pack2(q0[2], q0[3]), pack2(q0[4], q0[5]), pack2(q0[6], q0[7]));
        o0[1] = make_uint4(pack2(q0[8], q0[9]), pack2(q0[10], q0[11]), pack2(q0[12], q0[13]), pack2(q0[14], q0[15]));
        if (levels > 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) q1[i] = rh((q0[2 * i] + q0[2 * i + 1]) * 0.5f);
            *reinterpret_cast<uint4*>(v1 + row * Wl1 + j * 8) =
                make_uint4(pack2(q1[0], q1[1]), pack2(q1[2], q1[3]), pack2(q1[4], q1[5]), pack2(q1[6], q1[7]));
        }
        if (levels > 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) q2[i] = rh((q1[2 * i] + q1[2 * i + 1]) * 0.5f);
            *reinterpret_cast<uint2*>(v2 + row * Wl2 + j * 4) = make_uint2(pack2(q2[0], q2[1]), pack2(q2[2], q2[3]));
        }
        if (levels > 3) {
            q3[0] = rh((q2[0] + q2[1]) * 0.5f);
            q3[1] = rh((q2[2] + q2[3]) * 0.5f);
            *reinterpret_cast<uint32_t*>(v3 + row * Wl3 + j * 2) = pack2(q3[0], q3[1]);
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// =====================================================================================================================
// Backward of the build on the tensor cores (fp16): per (b, h), with g = d(loss)/d(level-0 volume) [W1 x W2],
//   dF1[d, x] = sum_y g[x, y] F2[d, y] / sqrt(D)        A_MN = false:  M = x, N = d, K = y;  A = g rows (K-major)
//   dF2[d, y] = sum_x g[x, y] F1[d, x] / sqrt(D)        A_MN = true :  M = y, N = d, K = x;  A = g columns (MN-major)
// computed transposed (C^T[m, d]) so that M is the 128-wide volume axis and N = D (a multiple of 16 up to 256, which
// M = 64/128 UMMA shapes accept); B = the other feature map's [D x K] panel, K-contiguous (K-major) in both cases.
// Shared layout: the same 128-byte core matrices as the forward -- K-major cores hold 8 MN rows x 16 B of 8 K elements.
// Thread = one TMEM lane = one m; its 16-column batches are 16 different d planes of the output, so a warp stores 64
// contiguous bytes per (d, 32 m).
// =====================================================================================================================
template <bool A_MN>
__global__ void __launch_bounds__(kTcThreads) corr_build_bwd_tc_kernel(int D, int H, int W1, int W2,
                                                                       const __half* __restrict__ fmap,
                                                                       const __half* __restrict__ g,
                                                                       __half* __restrict__ dfmap, float div,
                                                                       uint32_t tmem_cols) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y / H, h = blockIdx.y % H;
    const int Kdim = A_MN ? W1 : W2, Mdim = A_MN ? W2 : W1;
    const int KC = ((Kdim + 15) & ~15) >> 3;                    // K core groups, K padded to the UMMA K of 16 with zeros
    const int m_base = blockIdx.x * kTcM;
    const int NB = D >> 3;
    uint8_t* sA = smem;                                         // [KC][16 M cores][128 B]
    uint8_t* sB = smem + (size_t)KC * 2048;                     // [KC][NB N cores][128 B]
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    const __half* gbh = g + ((size_t)b * H + h) * W1 * W2;
    const int cl = lane >> 3, dl = lane & 7;
    const int KCg = (KC + 3) >> 2;
    if (A_MN) {      // g[x][y]: k = x (row), MN cores = chunks of 8 y
        for (int it = warp; it < KC * 4; it += kTcThreads / 32) {
            const int kc = it >> 2, c = (it & 3) * 4 + cl;
            const int x = kc * 8 + dl;
            const bool in = x < W1 && c * 8 < W2;
            cp_async16(sA + (size_t)kc * 2048 + c * 128 + dl * 16, in ? gbh + (size_t)x * W2 + c * 8 : gbh, in ? 16u : 0u);
        }
    } else {         // g[x][y]: m = x (row), K cores = chunks of 8 y
        for (int it = warp; it < 16 * KCg; it += kTcThreads / 32) {
            const int m8 = it / KCg, kc = (it % KCg) * 4 + cl;
            const int x = m_base + m8 * 8 + dl;
            if (kc < KC) {
                const bool in = x < W1 && kc * 8 < W2;
                cp_async16(sA + (size_t)kc * 2048 + m8 * 128 + dl * 16, in ? gbh + (size_t)x * W2 + kc * 8 : gbh, in ? 16u : 0u);
            }
        }
    }
    const size_t plane = (size_t)H * Kdim;
    const __half* fb = fmap + (size_t)b * D * plane + (size_t)h * Kdim;     // F[d][k], k contiguous
    for (int it = warp; it < NB * KCg; it += kTcThreads / 32) {
        const int d8 = it / KCg, kc = (it % KCg) * 4 + cl;
        if (kc < KC) {
            const bool in = kc * 8 < Kdim;
            cp_async16(sB + ((size_t)kc * NB + d8) * 128 + dl * 16, in ? fb + (size_t)(d8 * 8 + dl) * plane + kc * 8 : fb,
                       in ? 16u : 0u);
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (A_MN ? (1u << 15) : 0u) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        const uint32_t lbo_b = (uint32_t)NB * 128u;
        for (int s = 0; s < (KC >> 1); ++s) {
            const uint64_t da = umma_desc(a0 + (uint32_t)s * 2u * 2048u, 2048u, 128u);
            const uint64_t db = umma_desc(b0 + (uint32_t)s * 2u * lbo_b, lbo_b, 128u);
            umma_f16(tmem, da, db, idesc, s > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                     : "memory");
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int quad = warp & 3;
    const int m = m_base + quad * 32 + lane;
    const float rdiv = __frcp_rn(div);
    __half* o = dfmap + ((size_t)b * D * H + h) * Mdim + m;     // + d * H * Mdim
    const size_t dplane = (size_t)H * Mdim;
    for (int j = warp >> 2; j < (D >> 4); j += 2) {
        float acc[16];
        __syncwarp();
        tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * 16), acc);
        if (m < Mdim) {
#pragma unroll
            for (int i = 0; i < 16; ++i) o[(size_t)(j * 16 + i) * dplane] = __float2half_rn(div_rn_fast(acc[i], div, rdiv));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// true when the tensor-core kernel covers this problem (otherwise the caller uses the FFMA kernel)
bool corr_build_tc_supported(int dtype, int D, int W1, int W2, const void* f1, const void* f2, void* const* v, int levels) {
    if (dtype != 1) return false;
    if (D < 16 || (D & 15) || W1 < 8 || (W1 & 7) || W2 < 16 || (W2 & 15) || W2 > 128) return false;
    if ((size_t)D * 256 + (size_t)D * W2 * 2 > 200 * 1024) return false;
    if (!aligned16(f1) || !aligned16(f2)) return false;
    for (int l = 0; l < levels && l < 4; ++l)
        if (!aligned16(v[l])) return false;
    return true;
}

int launch_corr_build_tc(int B, int D, int H, int W1, int W2, const void* f1, const void* f2, void* v0, void* v1, void* v2,
                         void* v3, int levels, cudaStream_t stream) {
    const size_t smem = (size_t)D * 256 + (size_t)D * W2 * 2;
    GPSG_CUDA(cudaFuncSetAttribute(corr_build_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((W1 + kTcM - 1) / kTcM, B * H);
    GPSG_REQUIRE(grid.y <= 65535, "corr build: B*H too large");
    corr_build_tc_kernel<<<grid, kTcThreads, smem, stream>>>(D, H, W1, W2, (const __half*)f1, (const __half*)f2, (__half*)v0,
                                                            (__half*)v1, (__half*)v2, (__half*)v3, levels,
                                                            sqrtf((float)D));
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

bool corr_build_bwd_tc_supported(int dtype, int D, int W1, int W2, const void* f1, const void* f2, const void* g,
                                 const void* d1, const void* d2) {
    if (dtype != 1) return false;
    if (D < 16 || (D & 15) || D > 256 || W1 < 8 || (W1 & 7) || W2 < 16 || (W2 & 15) || W2 > 128) return false;
    const size_t kc1 = (size_t)((W2 + 15) & ~15) >> 3, kc2 = (size_t)((W1 + 15) & ~15) >> 3;
    if (kc1 * (2048 + (size_t)D * 16) > 200 * 1024 || kc2 * (2048 + (size_t)D * 16) > 200 * 1024) return false;
    return aligned16(f1) && aligned16(f2) && aligned16(g) && aligned16(d1) && aligned16(d2);
}

int launch_corr_build_bwd_tc(int B, int D, int H, int W1, int W2, const void* f1, const void* f2, const void* g, void* df1,
                             void* df2, cudaStream_t stream) {
    const uint32_t cols = D <= 32 ? 32u : D <= 64 ? 64u : D <= 128 ? 128u : 256u;
    const float div = sqrtf((float)D);
    const size_t kc1 = (size_t)((W2 + 15) & ~15) >> 3, kc2 = (size_t)((W1 + 15) & ~15) >> 3;
    const size_t smem1 = kc1 * (2048 + (size_t)D * 16), smem2 = kc2 * (2048 + (size_t)D * 16);
    GPSG_REQUIRE((size_t)B * H <= 65535, "corr build backward: B*H too large");
    GPSG_CUDA(cudaFuncSetAttribute(corr_build_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    GPSG_CUDA(cudaFuncSetAttribute(corr_build_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    corr_build_bwd_tc_kernel<false><<<dim3((W1 + kTcM - 1) / kTcM, B * H), kTcThreads, smem1, stream>>>(
        D, H, W1, W2, (const __half*)f2, (const __half*)g, (__half*)df1, div, cols);
    GPSG_LAUNCH_CHECK();
    corr_build_bwd_tc_kernel<true><<<dim3(1, B * H), kTcThreads, smem2, stream>>>(D, H, W1, W2, (const __half*)f1,
                                                                                (const __half*)g, (__half*)df2, div, cols);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

}  // namespace gpsg
