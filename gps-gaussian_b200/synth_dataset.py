"""Synthetic stereo-human dataset on disk in the layout the reference loader reads (harness item).

    <root>/<phase>/img/<sample>/<view>.jpg, <view>_hr.jpg      rgb, S x S and 2S x 2S     (lib/human_loader.py:109-110)
    <root>/<phase>/mask/<sample>/<view>.png                    3-channel 0/255             (:111, :287 uses channel 0)
    <root>/<phase>/depth/<sample>/<view>.png                   uint16 = INVERSE depth * 2^15, 0 = background
                                                               (:112, read_depth :93-94, depth2pts :28-49)
    <root>/<phase>/parm/<sample>/<view>_intrinsic.npy [3,3], <view>_extrinsic.npy [3,4]       (:113-114)

The subject is the analytic capsule of synth.py seen from the ring cameras of prepare_data/render_data.py (16 cameras on a
circle: source views 0/1 are 22.5 deg apart, novel views 2..4 sit between them); the texture is a smooth function of the
3-D surface point, so all views of a sample are photo-consistent.  Written with the same encodings as
prepare_data/render_data.py:11-32 (`save`).  Host-side numpy / OpenCV only; nothing here is on the hot path.
"""
import os

import cv2
import numpy as np

from . import synth

SOURCE_ANGLES = (-11.25, 11.25)                  # views 0, 1
NOVEL_ANGLES = (-5.625, 0.0, 5.625)              # views 2, 3, 4 (train_novel_id of config/stage2.yaml:12)


def view_angle(view_id):
    return (SOURCE_ANGLES + NOVEL_ANGLES)[view_id]


def _texture(xyz, phase):
    """Smooth photo-consistent colour in [0.1, 0.9] from world-space surface points [...,3]."""
    k = np.array([[9.0, 5.0, 7.0], [4.0, 11.0, 6.0], [7.0, 3.0, 10.0]])
    return 0.5 + 0.4 * np.sin(xyz @ k.T + phase)


def render_view(res, angle_deg, body_radius, phase):
    """One view of the capsule: (img float [res,res,3] in [0,1], mask [res,res] bool, z-depth [res,res], K, E)."""
    K, E = synth.ring_camera(angle_deg, res)
    z = synth._capsule_depth(K, E, res, body_radius, body_radius, 1.8 - body_radius)
    hit = z > 0
    v, u = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
    pc = np.stack([(u - K[0, 2]) * z / K[0, 0], (v - K[1, 2]) * z / K[1, 1], z], -1)
    xyz = (pc - E[:, 3]) @ E[:, :3]                                   # R^T (p - t)
    img = np.where(hit[..., None], _texture(xyz, phase), 0.0)
    return img, hit, z, K, E


def write_sample(root, phase, sample, res, body_radius=0.425, views=(0, 1, 2, 3, 4), hr=True, tex_phase=0.0):
    base = os.path.join(root, phase)
    for sub in ("img", "mask", "depth", "parm"):
        os.makedirs(os.path.join(base, sub, sample), exist_ok=True)
    to_u8 = lambda a: (np.clip(a, 0, 1) * 255.0 + 0.5).astype(np.uint8)
    for vid in views:
        img, hit, z, K, E = render_view(res, view_angle(vid), body_radius, tex_phase)
        inv = np.where(hit, 1.0 / np.maximum(z, 1e-6), 0.0)
        cv2.imwrite(os.path.join(base, "depth", sample, f"{vid}.png"), (inv * 2.0 ** 15).astype(np.uint16))
        cv2.imwrite(os.path.join(base, "img", sample, f"{vid}.jpg"), to_u8(img)[:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, 98])
        if hr:
            img_hr = render_view(2 * res, view_angle(vid), body_radius, tex_phase)[0]
            cv2.imwrite(os.path.join(base, "img", sample, f"{vid}_hr.jpg"), to_u8(img_hr)[:, :, ::-1],
                        [cv2.IMWRITE_JPEG_QUALITY, 98])
        cv2.imwrite(os.path.join(base, "mask", sample, f"{vid}.png"), np.repeat(to_u8(hit.astype(np.float64))[..., None], 3, -1))
        np.save(os.path.join(base, "parm", sample, f"{vid}_intrinsic.npy"), K.astype(np.float64))      # float64 like render_data.py (float32 scalars break graphics_utils.py:41 on new torch)
        np.save(os.path.join(base, "parm", sample, f"{vid}_extrinsic.npy"), E.astype(np.float64))


def write_dataset(root, n_train=2, n_val=1, res=256, body_radius=0.425, hr=True, seed=synth.SEED):
    """<root>/train and <root>/val with `n_*` samples each (different body radius / texture per sample)."""
    rng = np.random.default_rng(seed)
    out = {}
    for phase, n in (("train", n_train), ("val", n_val)):
        names = []
        for k in range(n):
            name = "%04d_%03d" % (k, 0)
            write_sample(root, phase, name, res, body_radius * float(rng.uniform(0.9, 1.05)), hr=hr,
                         tex_phase=float(rng.uniform(0, 6.28)))
            names.append(name)
        out[phase] = names
    return out
