"""Shared helpers for the parity tests (oracle = checker only)."""
import numpy as np

from oracle.raster_oracle import RasterOracle


def oracle_forward(sc, dtype="f32", nthreads=8, render=True):
    o = RasterOracle(dtype)
    return o, o.forward(sc["means3D"], sc["colors"], sc["opacity"], sc.get("scales"), sc.get("rots"), sc["view"],
                        sc["proj"], sc["tanfovx"], sc["tanfovy"], sc["W"], sc["H"], sc["bg"],
                        scale_mod=sc.get("scale_modifier", 1.0), cov3D_precomp=sc.get("cov3D_precomp"),
                        nthreads=nthreads, render=render)


def rel_err(a, b):
    """max |a-b| / max|b|  (normalised max error; b is the reference)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.abs(a - b).max() if a.size else 0.0
    s = np.abs(b).max() if b.size else 0.0
    return d / s if s > 0 else d
