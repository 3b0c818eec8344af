"""TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's cpu_baseline leg, never by the product).

CPU restatement of the photometric loss of reference lib/loss.py:35-72 + train_stage2.py:70-72, float64 numpy:
  l1   = mean |x - y|                                                    (lib/loss.py:35-36)
  ssim = mean over B*C*H*W of ((2 mu1 mu2 + C1)(2 s12 + C2)) / ((mu1^2 + mu2^2 + C1)(s1 + s2 + C2))   (:57-72)
  with mu = depthwise 11x11 Gaussian (sigma 1.5, float32 weights normalised in float32, :39-48) zero-padded conv.
PINNED by tests/golden/loss_golden.npz (values and autograd gradients of the reference's own functions)."""
import math

import numpy as np

C1, C2 = 0.01 ** 2, 0.03 ** 2


# float32 bit patterns of the reference's `gaussian(11, 1.5)` (lib/loss.py:39-41) as torch produces them (its float32
# sum/divide differs from numpy's by one ulp, which is visible in low-variance regions); checked against the window
# stored in tests/golden/loss_golden.npz by tests/test_loss.py.
_W_BITS = (981912246, 1006173953, 1024685452, 1038088319, 1046093343, 1049113264)


def window_1d():
    half = np.array(_W_BITS, np.uint32).view(np.float32)
    return np.concatenate([half, half[4::-1]]).astype(np.float32)


def _blur(a, w2d):
    """zero-padded 11x11 correlation of every [H,W] plane of a[...,H,W] (explicit window sum, no separability assumed)."""
    H, W = a.shape[-2:]
    p = np.zeros(a.shape[:-2] + (H + 10, W + 10))
    p[..., 5:5 + H, 5:5 + W] = a
    out = np.zeros_like(a, dtype=np.float64)
    for i in range(11):
        for j in range(11):
            out += w2d[i, j] * p[..., i:i + H, j:j + W]
    return out


def ssim_map(x, y):
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    w = window_1d()
    w2d = np.outer(w, w).astype(np.float32).astype(np.float64)           # float32 outer product, like _1D.mm(_1D.t()).float()
    mu1, mu2 = _blur(x, w2d), _blur(y, w2d)
    s1 = _blur(x * x, w2d) - mu1 * mu1
    s2 = _blur(y * y, w2d) - mu2 * mu2
    s12 = _blur(x * y, w2d) - mu1 * mu2
    return ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))


def l1(x, y):
    return float(np.abs(np.asarray(x, np.float64) - np.asarray(y, np.float64)).mean())


def ssim(x, y):
    return float(ssim_map(x, y).mean())


def loss(x, y, w_l1=0.8, w_ssim=0.2):
    return w_l1 * l1(x, y) + w_ssim * (1.0 - ssim(x, y))


def loss_grad_fd(x, y, idx, w_l1=0.8, w_ssim=0.2, eps=1e-6):
    """central finite difference of loss() w.r.t. x[idx] (spot checks of the analytic backward)."""
    x = np.array(x, np.float64)
    xp, xm = x.copy(), x.copy()
    xp[idx] += eps
    xm[idx] -= eps
    return (loss(xp, y, w_l1, w_ssim) - loss(xm, y, w_l1, w_ssim)) / (2 * eps)
