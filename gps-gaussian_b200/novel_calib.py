"""Novel-view calibration, batched and closed-form: mirror of reference lib/utils.py:8-54 (`get_novel_calib`).

The reference loops over the batch in Python and, PER NOVEL VIEW, pulls four calibration tensors to the host
(`.cpu().numpy()`, 4 syncs per sample), runs scipy `Slerp`, two 4x4 matrix inversions, a torch `bmm` and an
`inverse()`, and pushes five tensors back with `.cuda()`.  Here the same quantities come from closed forms, vectorised
over samples AND interpolation ratios in one numpy pass on the host (the rasterizer takes its camera BY VALUE in
`GpsgRasterSettings`, so nothing ever needs to go to the device):

  * rotation: quaternion slerp along the shorter arc == scipy `Slerp([0,1], [R0,R1])(ratio)` (lib/utils.py:20-24);
  * translation / intrinsics: lerp (lib/utils.py:28,30), `intr[:2] *= 2` when `use_hr_img` (:32-33);
  * world->view with `trans`/`scale`: camera centre c = -R^T t, c' = (c + trans)*scale, t' = -R c'
    == the inv/shift/inv of `getWorld2View2` (lib/graphics_utils.py:17-28);
  * projection: the off-centre matrix of `getProjectionMatrix` (lib/graphics_utils.py:31-48) simplified
    (2fx/w, 2fy/h, (2cx-w)/w, (2cy-h)/h, zfar/(zfar-znear), -zfar*znear/(zfar-znear));
  * camera centre = c' (== `world_view_transform.inverse()[3,:3]`, :41); FoV = 2 atan(pixels / 2 focal) (:36-37).

All outputs are float32 in the reference's row-vector (transposed) storage.  Pinned against the reference's own
`get_novel_calib` by tests/golden/novel_calib_golden.npz (tests/test_novel_views.py).
"""
import numpy as np
import torch


def _mat_to_quat(R):
    """[...,3,3] rotation matrices -> unit quaternions (w,x,y,z), numerically robust branch selection."""
    R = np.asarray(R, np.float64)
    m00, m01, m02 = R[..., 0, 0], R[..., 0, 1], R[..., 0, 2]
    m10, m11, m12 = R[..., 1, 0], R[..., 1, 1], R[..., 1, 2]
    m20, m21, m22 = R[..., 2, 0], R[..., 2, 1], R[..., 2, 2]
    cand = np.stack([
        np.stack([1 + m00 + m11 + m22, m21 - m12, m02 - m20, m10 - m01], -1),
        np.stack([m21 - m12, 1 + m00 - m11 - m22, m01 + m10, m02 + m20], -1),
        np.stack([m02 - m20, m01 + m10, 1 - m00 + m11 - m22, m12 + m21], -1),
        np.stack([m10 - m01, m02 + m20, m12 + m21, 1 - m00 - m11 + m22], -1)], -2)       # [...,4 branches,4]
    diag = np.stack([cand[..., 0, 0], cand[..., 1, 1], cand[..., 2, 2], cand[..., 3, 3]], -1)
    best = np.argmax(diag, -1)
    q = np.take_along_axis(cand, best[..., None, None], -2)[..., 0, :]
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def _quat_to_mat(q):
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
        np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)


def _slerp(R0, R1, ratios):
    """R0,R1 [bs,3,3]; ratios [n] -> [bs,n,3,3], shorter-arc geodesic from R0 (ratio 0) to R1 (ratio 1)."""
    q0, q1 = _mat_to_quat(R0), _mat_to_quat(R1)
    d = np.sum(q0 * q1, -1, keepdims=True)
    q1 = np.where(d < 0, -q1, q1)
    d = np.clip(np.abs(d), 0.0, 1.0)
    cross = np.linalg.norm(q1 - d * q0, axis=-1, keepdims=True)          # sin(theta), stable near 0
    theta = np.arctan2(cross, d)                                          # half the rotation angle, in [0, pi/2]
    r = np.asarray(ratios, np.float64)[None, :, None]
    th, s = theta[:, None, :], np.sin(theta)[:, None, :]
    small = s < 1e-8
    s_safe = np.where(small, 1.0, s)
    w0 = np.where(small, 1.0 - r, np.sin((1.0 - r) * th) / s_safe)
    w1 = np.where(small, r, np.sin(r * th) / s_safe)
    q = w0 * q0[:, None, :] + w1 * q1[:, None, :]
    return _quat_to_mat(q / np.linalg.norm(q, axis=-1, keepdims=True))


def novel_calib_batch(intr0, intr1, extr0, extr1, ratios, width, height, znear, zfar, trans=(0.0, 0.0, 0.0), scale=1.0,
                      use_hr_img=False):
    """intr* [bs,3,3], extr* [bs,3,4] (lmain / rmain), ratios [n], width/height scalars or [bs].
    Returns float32 numpy arrays: FovX, FovY [bs,n]; world_view_transform, full_proj_transform [bs,n,4,4];
    camera_center [bs,n,3] -- the values lib/utils.py:8-54 stores in data['novel_view'] for each ratio."""
    K0, K1 = np.asarray(intr0, np.float64), np.asarray(intr1, np.float64)
    E0, E1 = np.asarray(extr0, np.float64), np.asarray(extr1, np.float64)
    bs = K0.shape[0]
    r = np.asarray(ratios, np.float64).reshape(-1)
    n = r.shape[0]
    w = np.broadcast_to(np.asarray(width, np.float64).reshape(-1), (bs,))[:, None]
    h = np.broadcast_to(np.asarray(height, np.float64).reshape(-1), (bs,))[:, None]
    R = _slerp(E0[:, :3, :3], E1[:, :3, :3], r).astype(np.float32).astype(np.float64)      # npose is float32 (:25-27)
    t = ((1.0 - r)[None, :, None] * E0[:, None, :3, 3] + r[None, :, None] * E1[:, None, :3, 3]).astype(np.float32).astype(np.float64)
    K = (1.0 - r)[None, :, None, None] * K0[:, None] + r[None, :, None, None] * K1[:, None]
    if use_hr_img:
        K = K.copy()
        K[:, :, :2] *= 2
    c = -np.einsum("bnji,bnj->bni", R, t)                                   # camera centre  -R^T t
    c = (c + np.asarray(trans, np.float64)) * float(scale)
    t2 = -np.einsum("bnij,bnj->bni", R, c)
    w2v = np.zeros((bs, n, 4, 4))
    w2v[..., :3, :3], w2v[..., :3, 3], w2v[..., 3, 3] = R, t2, 1.0
    wvt = np.swapaxes(w2v, -1, -2).astype(np.float32)
    fx, fy, cx, cy = K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2]
    P = np.zeros((bs, n, 4, 4))
    P[..., 0, 0], P[..., 1, 1] = 2.0 * fx / w, 2.0 * fy / h
    P[..., 0, 2], P[..., 1, 2] = (2.0 * cx - w) / w, (2.0 * cy - h) / h
    P[..., 3, 2] = 1.0
    P[..., 2, 2], P[..., 2, 3] = zfar / (zfar - znear), -(zfar * znear) / (zfar - znear)
    full = (wvt.astype(np.float64) @ np.swapaxes(P.astype(np.float32).astype(np.float64), -1, -2)).astype(np.float32)
    return dict(FovX=(2.0 * np.arctan(w / (2.0 * fx))).astype(np.float32), FovY=(2.0 * np.arctan(h / (2.0 * fy))).astype(np.float32),
                world_view_transform=wvt, full_proj_transform=full, camera_center=c.astype(np.float32))


def _host(t):
    return t.detach().to("cpu", torch.float64).numpy() if isinstance(t, torch.Tensor) else np.asarray(t, np.float64)


def calib_from_data(data, opt, ratios, intr_key='intr', extr_key='extr'):
    """One host read of the pair's calibration (2 small D2H copies if it lives on the GPU), all ratios at once."""
    nv = data['novel_view']
    return novel_calib_batch(_host(data['lmain'][intr_key]), _host(data['rmain'][intr_key]), _host(data['lmain'][extr_key]),
                             _host(data['rmain'][extr_key]), ratios, _host(nv['width']), _host(nv['height']),
                             opt.znear, opt.zfar, getattr(opt, 'trans', (0.0, 0.0, 0.0)), getattr(opt, 'scale', 1.0),
                             bool(getattr(opt, 'use_hr_img', False)))


def get_novel_calib(data, opt, ratio=0.5, intr_key='intr', extr_key='extr'):
    """Same signature and effect as reference lib/utils.py:8-54: fills data['novel_view'][FovX, FovY,
    world_view_transform, full_proj_transform, camera_center] for one ratio.  The tensors stay on the host (pinned when
    CUDA is present): `render` reads the camera by value, so the reference's five `.cuda()` copies are not needed --
    it accepts host camera tensors just as it does in training (train_stage2.py:155-157)."""
    c = calib_from_data(data, opt, [ratio], intr_key, extr_key)
    nv = data['novel_view']
    for k, v in c.items():
        t = torch.from_numpy(np.ascontiguousarray(v[:, 0]))
        nv[k] = t.pin_memory() if torch.cuda.is_available() else t
    return data
