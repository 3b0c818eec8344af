/*
 * taichi_splat_oracle.c -- CPU BASELINE restatement (test/bench infrastructure, NOT product code).
 *
 * Restates the Taichi kernel `TaichiRenderBatch.render_respective_color`
 * (reference lib/TaichiRender.py:12-23): a 1-pixel z-buffer point splat using
 * atomic_max on INVERSE depth.  It is NOT a Gaussian rasterizer (different
 * algorithm, no parity claim); BASELINE.json names it as the reference's
 * "CPU path", so bench.py reports its host-core throughput beside the GPU number.
 * `taichi==1.5.0` (reference environment.yml:17) is not installed here, hence the port.
 *
 * Sequential semantics (race-free reading of the Taichi struct-for): points in index order.
 *   pts[B,N,6] = (x_pix, y_pix, inv_depth, r, g, b), mask[B,N], depth[B,res,res], color[B,3,res,res]
 */
#include "oracle_real.h"

void SUFFIX(oracle_taichi_splat)(int B, int N, int res, const real* pts, const real* mask, real* depth, real* color) {
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            if (mask[(size_t)b * N + i] < RC(0.5)) continue;
            const real* p = pts + ((size_t)b * N + i) * 6;
            int IX = (int)p[0], IY = (int)p[1]; /* ti.cast(f32->i32) truncates */
            IX = IX < 0 ? 0 : (IX > res - 1 ? res - 1 : IX);
            IY = IY < 0 ? 0 : (IY > res - 1 ? res - 1 : IY);
            real* d = depth + ((size_t)b * res + IY) * res + IX;
            real old = *d;
            if (p[2] > old) *d = p[2]; /* atomic_max returns the OLD value */
            if (p[2] >= old)
                for (int k = 0; k < 3; ++k) color[(((size_t)b * 3 + k) * res + IY) * res + IX] = p[3 + k];
        }
}
