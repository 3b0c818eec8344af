"""ctypes front-end of the CPU correlation oracle (oracle/corr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Restates reference core/corr.py:31-61 (`CorrBlockFast1D`: volume, pyramid, 4-level radius-r lookup) with the
`corr_sampler` semantics of SURVEY.md Appendix B.  Pinned by tests/golden/corr_golden.npz, which was produced by
the reference's own PyTorch twin `CorrBlock1D` (core/corr.py:110-156) -- see tests/golden/make_corr_golden.py."""
import ctypes as C

import numpy as np

from .raster_oracle import _lib, _p


class CorrOracle:
    def __init__(self, dtype="f32"):
        self.tag = dtype
        self.np = np.float32 if dtype == "f32" else np.float64
        self.lib = _lib(dtype)

    def _fn(self, n):
        return getattr(self.lib, f"{n}_{self.tag}")

    def _a(self, x):
        return np.ascontiguousarray(np.asarray(x, dtype=self.np))

    def volume(self, f1, f2):
        f1, f2 = self._a(f1), self._a(f2)
        B, D, H, W1 = f1.shape
        W2 = f2.shape[3]
        vol = np.zeros((B, H, W1, W2), self.np)
        self._fn("oracle_corr_volume")(C.c_int(B), C.c_int(D), C.c_int(H), C.c_int(W1), C.c_int(W2), _p(f1), _p(f2), _p(vol))
        return vol

    def pool(self, vol):
        vol = self._a(vol)
        W = vol.shape[-1]
        rows = int(np.prod(vol.shape[:-1]))
        out = np.zeros(vol.shape[:-1] + (W // 2,), self.np)
        self._fn("oracle_corr_pool")(C.c_int64(rows), C.c_int(W), _p(vol), _p(out))
        return out

    def pyramid(self, f1, f2, num_levels=4):
        lv = [self.volume(f1, f2)]
        for _ in range(num_levels - 1):
            lv.append(self.pool(lv[-1]))
        return lv

    def sample_fwd(self, vol, coords_x, r):
        vol, c = self._a(vol), self._a(coords_x)
        B, H, W1, W2 = vol.shape
        out = np.zeros((B, 2 * r + 1, H, W1), self.np)
        self._fn("oracle_corr_sample_fwd")(C.c_int(B), C.c_int(H), C.c_int(W1), C.c_int(W2), _p(vol), _p(c), C.c_int(r), _p(out))
        return out

    def sample_bwd(self, shape, coords_x, grad_out, r):
        B, H, W1, W2 = shape
        c, g = self._a(coords_x), self._a(grad_out)
        gv = np.zeros((B, H, W1, W2), self.np)
        self._fn("oracle_corr_sample_bwd")(C.c_int(B), C.c_int(H), C.c_int(W1), C.c_int(W2), _p(c), _p(g), C.c_int(r), _p(gv))
        return gv

    def lookup(self, pyramid, coords, r=4):
        """CorrBlockFast1D.__call__ (core/corr.py:44-51): coords [B,2,H,W] -> [B, levels*(2r+1), H, W]."""
        cx = np.asarray(coords)[:, 0]
        return np.concatenate([self.sample_fwd(v, cx / self.np(2 ** i), r) for i, v in enumerate(pyramid)], 1)
