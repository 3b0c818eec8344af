"""Second, fully independent fp64 restatement of the rasterizer FORWARD (TEST INFRASTRUCTURE ONLY).

VERDICT r1 "missing" item 6: `oracle/raster_torch64.py` borrows the C oracle's binning state, so it cannot catch an error
that `oracle/gpsg_oracle.c` makes in culling, radius, tile rectangles, key order or tile ranges.  This module takes ONLY
the raw call arguments of `rasterize_gaussians` (reference call site gaussian_renderer/__init__.py:36-62) and restates
SURVEY.md Appendix A.2-A.4 in numpy from the maths form -- matrix products instead of the spelled-out scalar chains, a
global stable argsort on the 64-bit (tile, depth) key instead of per-tile qsort, vectorised per-tile compositing with
cumulative products instead of a per-pixel loop.  It shares no code and no state with gpsg_oracle.c / raster_oracle.py.
tests/test_oracle_cpu.py pins the two against each other (integer outputs identical, image to 1e-12);
tests/test_raster_gpu.py checks the device against this one directly.

Still "parity unpinned" in the sense of DESIGN.md section 2: both restate the same published algorithm; neither has
seen the real extension's output.  tools/dump_reference_vectors.py is the script to run where the real extension exists.
"""
import numpy as np

TILE = 16


def _quat_to_rot(q):
    r, x, y, z = (q[:, k] for k in range(4))                       # (r, x, y, z), NOT normalised (A.2 step 3)
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _trunc_div16(v):
    """(int)(v / 16) with C truncation toward zero."""
    return np.trunc(v / TILE).astype(np.int64)


def project(means3D, scales, rots, opacity, view, proj, tanfovx, tanfovy, W, H, scale_mod=1.0):
    """A.2.  view/proj: the 4x4 tensors as the reference passes them (row-vector convention: p_row @ M).
    Returns dict(radii, pix[P,2], depth, conic[P,3], rect[P,4] = (x0, y0, x1, y1) tile bounds, visible)."""
    f8 = np.float64
    m = np.asarray(means3D, f8).reshape(-1, 3)
    P = m.shape[0]
    V = np.asarray(view, f8).reshape(4, 4)
    Pm = np.asarray(proj, f8).reshape(4, 4)
    hom = np.concatenate([m, np.ones((P, 1))], 1)
    t = (hom @ V)[:, :3]
    ph = hom @ Pm
    w_inv = 1.0 / (ph[:, 3] + f8(np.float32(1e-7)))
    ndc = ph[:, :2] * w_inv[:, None]
    in_front = t[:, 2] > f8(np.float32(0.2))
    R = _quat_to_rot(np.asarray(rots, f8).reshape(-1, 4))
    s = np.asarray(scales, f8).reshape(-1, 3) * scale_mod
    Sigma = np.einsum("pik,pk,pjk->pij", R, s * s, R)               # R diag(s^2) R^T
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    limx, limy = f8(np.float32(1.3)) * tanfovx, f8(np.float32(1.3)) * tanfovy
    tz = np.where(in_front, t[:, 2], 1.0)
    tx = np.clip(t[:, 0] / tz, -limx, limx) * tz
    ty = np.clip(t[:, 1] / tz, -limy, limy) * tz
    J = np.zeros((P, 2, 3))
    J[:, 0, 0] = fx / tz; J[:, 0, 2] = -fx * tx / (tz * tz)
    J[:, 1, 1] = fy / tz; J[:, 1, 2] = -fy * ty / (tz * tz)
    A = J @ V[:3, :3].T                                             # V[:3,:3].T = maths world->view rotation
    cov = A @ Sigma @ np.swapaxes(A, 1, 2)
    k03 = f8(np.float32(0.3))
    a, b, c = cov[:, 0, 0] + k03, cov[:, 0, 1], cov[:, 1, 1] + k03
    det = a * c - b * b
    ok = in_front & (det != 0)
    det_s = np.where(ok, det, 1.0)
    conic = np.stack([c / det_s, -b / det_s, a / det_s], 1)
    mid = 0.5 * (a + c)
    lam = mid + np.sqrt(np.maximum(f8(np.float32(0.1)), mid * mid - det))
    with np.errstate(invalid="ignore"):
        rad_f = np.ceil(3.0 * np.sqrt(lam))
    ok &= np.isfinite(rad_f) & (rad_f > 0)                         # NaN covariance -> culled (finite inputs: radius >= 2)
    radius = np.where(ok, rad_f, 0).astype(np.int64)
    pix = np.stack([((ndc[:, 0] + 1.0) * W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5], 1)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    rad = radius.astype(f8)
    x0 = np.clip(_trunc_div16(pix[:, 0] - rad), 0, gx); x1 = np.clip(_trunc_div16(pix[:, 0] + rad + (TILE - 1)), 0, gx)
    y0 = np.clip(_trunc_div16(pix[:, 1] - rad), 0, gy); y1 = np.clip(_trunc_div16(pix[:, 1] + rad + (TILE - 1)), 0, gy)
    visible = ok & ((x1 - x0) * (y1 - y0) > 0)
    z = lambda v: np.where(visible.reshape((-1,) + (1,) * (v.ndim - 1)), v, 0)
    return dict(radii=z(radius).astype(np.int32), pix=z(pix), depth=z(t[:, 2]), conic=z(conic),
                rect=z(np.stack([x0, y0, x1, y1], 1)).astype(np.int64), visible=visible, gx=gx, gy=gy)


def bin_tiles(g):
    """A.3: (tile, depth) keys in emission order, one global STABLE sort, tile ranges.  Returns keys, point_list, ranges."""
    vis = np.nonzero(g["visible"])[0]
    r = g["rect"][vis]
    nx, ny = r[:, 2] - r[:, 0], r[:, 3] - r[:, 1]
    cnt = nx * ny
    N = int(cnt.sum())
    T = g["gx"] * g["gy"]
    ranges = np.zeros((T, 2), np.uint32)
    if N == 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint32), ranges
    owner = np.repeat(np.arange(vis.size), cnt)                      # emission order: Gaussian index, then y, then x
    local = np.arange(N) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    ty = r[owner, 1] + local // nx[owner]
    tx = r[owner, 0] + local % nx[owner]
    depth_bits = g["depth"].astype(np.float32).view(np.uint32).astype(np.uint64)
    keys = ((ty * g["gx"] + tx).astype(np.uint64) << np.uint64(32)) | depth_bits[vis][owner]
    order = np.argsort(keys, kind="stable")
    keys, plist = keys[order], vis[owner][order].astype(np.uint32)
    tile = (keys >> np.uint64(32)).astype(np.int64)
    first = np.nonzero(np.r_[True, tile[1:] != tile[:-1]])[0]
    last = np.r_[first[1:], N]
    ranges[tile[first], 0] = first
    ranges[tile[first], 1] = last
    return keys, plist, ranges


def composite(g, plist, ranges, colors, opacity, bg, W, H):
    """A.4, tile by tile, all pixels x all list entries at once (cumulative products for the transmittance)."""
    col = np.asarray(colors, np.float64).reshape(-1, 3)
    op = np.asarray(opacity, np.float64).reshape(-1)
    bg = np.asarray(bg, np.float64).reshape(3)
    img = np.empty((3, H, W)); final_T = np.ones((H, W)); n_contrib = np.zeros((H, W), np.uint32)
    img[:] = bg[:, None, None]
    thr_a = np.float64(np.float32(1.0) / np.float32(255.0))
    thr_T = np.float64(np.float32(0.0001))
    a_max = np.float64(np.float32(0.99))
    for t in range(ranges.shape[0]):
        s, e = int(ranges[t, 0]), int(ranges[t, 1])
        if e <= s:
            continue
        y0, x0 = (t // g["gx"]) * TILE, (t % g["gx"]) * TILE
        ys, xs = np.arange(y0, min(y0 + TILE, H)), np.arange(x0, min(x0 + TILE, W))
        ids = plist[s:e].astype(np.int64)
        py, px = np.meshgrid(ys.astype(np.float64), xs.astype(np.float64), indexing="ij")
        dx = g["pix"][ids, 0][None, :] - px.reshape(-1, 1)
        dy = g["pix"][ids, 1][None, :] - py.reshape(-1, 1)
        cn = g["conic"][ids]
        power = -0.5 * (cn[:, 0] * dx * dx + cn[:, 2] * dy * dy) - cn[:, 1] * dx * dy
        alpha = np.minimum(a_max, op[ids] * np.exp(np.minimum(power, 0.0)))
        keep = (power <= 0) & (alpha >= thr_a)
        alpha = np.where(keep, alpha, 0.0)
        T_after = np.cumprod(1.0 - alpha, 1)
        stop = keep & (T_after < thr_T)
        dead = np.cumsum(stop, 1) > 0                                  # the stopping entry and everything behind it
        alpha = np.where(dead, 0.0, alpha)
        T_before = np.concatenate([np.ones((alpha.shape[0], 1)), np.cumprod(1.0 - alpha, 1)[:, :-1]], 1)
        wgt = alpha * T_before
        T_fin = T_before[:, -1] * (1.0 - alpha[:, -1])
        out = wgt @ col[ids] + T_fin[:, None] * bg[None]
        contributed = wgt > 0
        last = np.where(contributed.any(1), alpha.shape[1] - np.argmax(contributed[:, ::-1], 1), 0)
        sl = (slice(y0, y0 + len(ys)), slice(x0, x0 + len(xs)))
        img[(slice(None),) + sl] = out.T.reshape(3, len(ys), len(xs))
        final_T[sl] = T_fin.reshape(len(ys), len(xs))
        n_contrib[sl] = last.reshape(len(ys), len(xs))
    return img, final_T, n_contrib


def forward(means3D, colors, opacity, scales, rots, view, proj, tanfovx, tanfovy, W, H, bg, scale_mod=1.0):
    g = project(means3D, scales, rots, opacity, view, proj, tanfovx, tanfovy, W, H, scale_mod)
    keys, plist, ranges = bin_tiles(g)
    img, final_T, n_contrib = composite(g, plist, ranges, colors, opacity, bg, W, H)
    rect = g["rect"]
    return dict(radii=g["radii"], means2D=g["pix"], depth=g["depth"], conic=g["conic"],
                tiles_touched=np.where(g["visible"], (rect[:, 2] - rect[:, 0]) * (rect[:, 3] - rect[:, 1]), 0).astype(np.uint32),
                keys=keys, point_list=plist, ranges=ranges, num_rendered=int(keys.size), color=img, final_T=final_T,
                n_contrib=n_contrib)


def forward_scene(sc):
    return forward(sc["means3D"], sc["colors"], sc["opacity"], sc["scales"], sc["rots"], sc["view"], sc["proj"],
                   sc["tanfovx"], sc["tanfovy"], sc["W"], sc["H"], sc["bg"], scale_mod=sc.get("scale_modifier", 1.0))
