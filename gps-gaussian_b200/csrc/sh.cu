// sh.cu -- spherical-harmonics colour branch (upstream
// forward.cu / backward.cu ::computeColorFromSH).  GPS-Gaussian itself always passes colors_precomp
// (reference gaussian_renderer/__init__.py:57-58); this exists so that `GaussianRasterizer(shs=...)` works.
#include "gpsg_internal.cuh"

namespace gpsg {

__device__ __constant__ float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                        -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                        0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                        -0.5900435899266435f};
constexpr float kC0 = 0.28209479177387814f, kC1 = 0.4886025119029199f;

// basis b[k] (k < (deg+1)^2) and, if GRAD, its gradient wrt the unit direction
template <bool GRAD>
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b, float (*db)[3]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        b[k] = 0.f;
        if (GRAD) { db[k][0] = 0.f; db[k][1] = 0.f; db[k][2] = 0.f; }
    }
    b[0] = kC0;
    if (deg < 1) return;
    b[1] = -kC1 * y; b[2] = kC1 * z; b[3] = -kC1 * x;
    if (GRAD) { db[1][1] = -kC1; db[2][2] = kC1; db[3][0] = -kC1; }
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = kC2[0] * xy; b[5] = kC2[1] * yz; b[6] = kC2[2] * (2.f * zz - xx - yy); b[7] = kC2[3] * xz; b[8] = kC2[4] * (xx - yy);
    if (GRAD) {
        db[4][0] = kC2[0] * y; db[4][1] = kC2[0] * x;
        db[5][1] = kC2[1] * z; db[5][2] = kC2[1] * y;
        db[6][0] = kC2[2] * (-2.f * x); db[6][1] = kC2[2] * (-2.f * y); db[6][2] = kC2[2] * (4.f * z);
        db[7][0] = kC2[3] * z; db[7][2] = kC2[3] * x;
        db[8][0] = kC2[4] * (2.f * x); db[8][1] = kC2[4] * (-2.f * y);
    }
    if (deg < 3) return;
    b[9] = kC3[0] * y * (3.f * xx - yy);
    b[10] = kC3[1] * xy * z;
    b[11] = kC3[2] * y * (4.f * zz - xx - yy);
    b[12] = kC3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
    b[13] = kC3[4] * x * (4.f * zz - xx - yy);
    b[14] = kC3[5] * z * (xx - yy);
    b[15] = kC3[6] * x * (xx - 3.f * yy);
    if (GRAD) {
        db[9][0] = kC3[0] * 6.f * xy; db[9][1] = kC3[0] * (3.f * xx - 3.f * yy);
        db[10][0] = kC3[1] * yz; db[10][1] = kC3[1] * xz; db[10][2] = kC3[1] * xy;
        db[11][0] = kC3[2] * (-2.f * xy); db[11][1] = kC3[2] * (4.f * zz - xx - 3.f * yy); db[11][2] = kC3[2] * 8.f * yz;
        db[12][0] = kC3[3] * (-6.f * xz); db[12][1] = kC3[3] * (-6.f * yz); db[12][2] = kC3[3] * (6.f * zz - 3.f * xx - 3.f * yy);
        db[13][0] = kC3[4] * (4.f * zz - 3.f * xx - yy); db[13][1] = kC3[4] * (-2.f * xy); db[13][2] = kC3[4] * 8.f * xz;
        db[14][0] = kC3[5] * 2.f * xz; db[14][1] = kC3[5] * (-2.f * yz); db[14][2] = kC3[5] * (xx - yy);
        db[15][0] = kC3[6] * (3.f * xx - 3.f * yy); db[15][1] = kC3[6] * (-6.f * xy);
    }
}

struct Campos { float x, y, z; };

__global__ void __launch_bounds__(256) sh_forward_kernel(int P, int deg, int M, Campos cp,
                                                         const float* __restrict__ means3D,
                                                         const float* __restrict__ shs,
                                                         const int32_t* __restrict__ radii, float* __restrict__ rgb,
                                                         uint8_t* __restrict__ clamped) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || radii[i] <= 0) return;
    const float dx = means3D[3 * i] - cp.x, dy = means3D[3 * i + 1] - cp.y, dz = means3D[3 * i + 2] - cp.z;
    const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
    float b[16];
    sh_basis<false>(deg, dx * inv, dy * inv, dz * inv, b, nullptr);
    const int nb = (deg + 1) * (deg + 1);
    const float* sh = shs + (size_t)i * M * 3;
    float c[3] = {0.5f, 0.5f, 0.5f};
    for (int k = 0; k < nb; ++k) {
        c[0] = fmaf(b[k], sh[3 * k], c[0]);
        c[1] = fmaf(b[k], sh[3 * k + 1], c[1]);
        c[2] = fmaf(b[k], sh[3 * k + 2], c[2]);
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        clamped[3 * i + ch] = c[ch] < 0.f;
        rgb[3 * i + ch] = fmaxf(c[ch], 0.f);
    }
}

__global__ void __launch_bounds__(256) sh_backward_kernel(int P, int deg, int M, Campos cp,
                                                          const float* __restrict__ means3D,
                                                          const float* __restrict__ shs,
                                                          const int32_t* __restrict__ radii,
                                                          const uint8_t* __restrict__ clamped,
                                                          const float* __restrict__ dL_dcolors,
                                                          float* __restrict__ dL_dsh, float* __restrict__ dL_dmeans3D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float* dsh = dL_dsh + (size_t)i * M * 3;
    if (radii[i] <= 0) {
        for (int k = 0; k < 3 * M; ++k) dsh[k] = 0.f;
        return;
    }
    const float vx = means3D[3 * i] - cp.x, vy = means3D[3 * i + 1] - cp.y, vz = means3D[3 * i + 2] - cp.z;
    const float inv = rsqrtf(vx * vx + vy * vy + vz * vz);
    const float d[3] = {vx * inv, vy * inv, vz * inv};
    float b[16], db[16][3];
    sh_basis<true>(deg, d[0], d[1], d[2], b, db);
    float g[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) g[ch] = clamped[3 * i + ch] ? 0.f : dL_dcolors[3 * i + ch];
    const int nb = (deg + 1) * (deg + 1);
    const float* sh = shs + (size_t)i * M * 3;
    float ddir[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < M; ++k) {
        if (k < nb) {
            dsh[3 * k] = b[k] * g[0]; dsh[3 * k + 1] = b[k] * g[1]; dsh[3 * k + 2] = b[k] * g[2];
            const float dot = sh[3 * k] * g[0] + sh[3 * k + 1] * g[1] + sh[3 * k + 2] * g[2];
            ddir[0] = fmaf(db[k][0], dot, ddir[0]);
            ddir[1] = fmaf(db[k][1], dot, ddir[1]);
            ddir[2] = fmaf(db[k][2], dot, ddir[2]);
        } else {
            dsh[3 * k] = 0.f; dsh[3 * k + 1] = 0.f; dsh[3 * k + 2] = 0.f;
        }
    }
    const float dd = d[0] * ddir[0] + d[1] * ddir[1] + d[2] * ddir[2];   // through d = v/|v|
#pragma unroll
    for (int a = 0; a < 3; ++a) dL_dmeans3D[3 * i + a] += (ddir[a] - d[a] * dd) * inv;
}

int launch_sh_forward(int P, int deg, int M, const float* campos3, const float* means3D, const float* shs,
                      const int32_t* radii, float* rgb, uint8_t* clamped, cudaStream_t stream) {
    if (P <= 0) return GPSG_OK;
    Campos cp{campos3[0], campos3[1], campos3[2]};
    sh_forward_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, deg, M, cp, means3D, shs, radii, rgb, clamped);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

int launch_sh_backward(int P, int deg, int M, const float* campos3, const float* means3D, const float* shs,
                       const int32_t* radii, const uint8_t* clamped, const float* dL_dcolors, float* dL_dsh,
                       float* dL_dmeans3D, cudaStream_t stream) {
    if (P <= 0) return GPSG_OK;
    Campos cp{campos3[0], campos3[1], campos3[2]};
    sh_backward_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, deg, M, cp, means3D, shs, radii, clamped, dL_dcolors,
                                                           dL_dsh, dL_dmeans3D);
    GPSG_LAUNCH_CHECK();
    return GPSG_OK;
}

}  // namespace gpsg
