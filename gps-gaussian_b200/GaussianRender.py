"""Mirror of reference lib/GaussianRender.py:5-39 -- `pts2render(data, bg_color)`: per sample, take the
valid pixel-aligned Gaussians of both source views, render the novel view, stack into
data['novel_view']['img_pred'] [B,3,H,W]."""
import torch

from .gaussian_renderer import render


def pts2render(data, bg_color):
    bs = data['lmain']['img'].shape[0]
    out = []
    for i in range(bs):
        parts = {k: [] for k in ('xyz', 'rgb', 'rot', 'scale', 'opacity')}
        for view in ('lmain', 'rmain'):
            d = data[view]
            valid = d['pts_valid'][i, :]
            parts['xyz'].append(d['xyz'][i][valid].view(-1, 3))
            parts['rgb'].append(d['img'][i].permute(1, 2, 0).reshape(-1, 3)[valid].view(-1, 3))
            parts['rot'].append(d['rot_maps'][i].permute(1, 2, 0).reshape(-1, 4)[valid].view(-1, 4))
            parts['scale'].append(d['scale_maps'][i].permute(1, 2, 0).reshape(-1, 3)[valid].view(-1, 3))
            parts['opacity'].append(d['opacity_maps'][i].permute(1, 2, 0).reshape(-1, 1)[valid].view(-1, 1))
        xyz = torch.cat(parts['xyz'], 0)
        rgb = torch.cat(parts['rgb'], 0) * 0.5 + 0.5
        img = render(data, i, xyz, rgb, torch.cat(parts['rot'], 0), torch.cat(parts['scale'], 0),
                     torch.cat(parts['opacity'], 0), bg_color=bg_color)
        out.append(img.unsqueeze(0))
    data['novel_view']['img_pred'] = torch.cat(out, 0)
    return data
