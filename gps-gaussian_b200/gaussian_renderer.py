"""Mirror of reference gaussian_renderer/__init__.py:17-67 -- `render(data, idx, pts_xyz, pts_rgb,
rotations, scales, opacity, bg_color)` with the same argument meaning and return value
([3,H,W] image; radii dropped), wired to the sm_100a rasterizer.

Differences that are invisible to the caller: the zero `screenspace_points` tensor is only created
when a gradient is being recorded (it exists upstream solely so autograd can return a means2D
gradient), and FoV/size scalars are read without forcing a device sync when they live on the host
(they do in training: reference train_stage2.py:155-157).
"""
import math
import os
import sys

import torch

_DROPIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
if _DROPIN not in sys.path:
    sys.path.insert(0, _DROPIN)
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402


def render(data, idx, pts_xyz, pts_rgb, rotations, scales, opacity, bg_color):
    nv = data['novel_view']
    bg = torch.as_tensor(bg_color, dtype=torch.float32)          # kept on host: passed by value to the kernels
    screenspace_points = torch.zeros_like(pts_xyz, dtype=torch.float32, requires_grad=True) + 0 \
        if torch.is_grad_enabled() and pts_xyz.requires_grad else torch.zeros_like(pts_xyz, dtype=torch.float32)
    if screenspace_points.requires_grad:
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
    tanfovx = math.tan(float(nv['FovX'][idx]) * 0.5)
    tanfovy = math.tan(float(nv['FovY'][idx]) * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(nv['height'][idx]), image_width=int(nv['width'][idx]), tanfovx=tanfovx, tanfovy=tanfovy,
        bg=bg, scale_modifier=1.0, viewmatrix=nv['world_view_transform'][idx],
        projmatrix=nv['full_proj_transform'][idx], sh_degree=3, campos=nv['camera_center'][idx],
        prefiltered=False, debug=False)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    rendered_image, _ = rasterizer(means3D=pts_xyz, means2D=screenspace_points, shs=None, colors_precomp=pts_rgb,
                                   opacities=opacity, scales=scales, rotations=rotations, cov3D_precomp=None)
    return rendered_image
