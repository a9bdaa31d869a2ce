"""CPU restatements of the rows next to the hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/ imports this module.  Each function cites the reference lines it follows; all three are
PINNED by golden vectors produced by executing the reference's own Python (utils/loss_utils.py, the
in_frustum functions) or the library the reference calls (torch.optim.Adam) on the CPU:
tests/golden/make_golden_next.py -> tests/golden/{loss_l1_ssim,frustum,adam}.npz.
"""
from __future__ import annotations

import ctypes as C
from math import exp

import numpy as np

from . import oracle as O


# ------------------------------------------------------------------------------ L1 + SSIM
def gaussian_window(window_size=11, sigma=1.5):
    """/root/reference/utils/loss_utils.py:23-25: fp32 taps, normalised in fp32."""
    g = np.array([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)], dtype=np.float32)
    return g / g.sum(dtype=np.float32)


def _conv(img, win2d):
    """Zero-padded 'same' correlation of every plane with the 11x11 window (conv2d(padding=5, groups=C),
    loss_utils.py:45-54), accumulated in float64."""
    r = win2d.shape[0] // 2
    C_, H, W = img.shape
    pad = np.zeros((C_, H + 2 * r, W + 2 * r), dtype=np.float64)
    pad[:, r:r + H, r:r + W] = img
    out = np.zeros((C_, H, W), dtype=np.float64)
    for dy in range(2 * r + 1):
        for dx in range(2 * r + 1):
            out += win2d[dy, dx] * pad[:, dy:dy + H, dx:dx + W]
    return out


def l1_ssim(img, gt):
    """loss_utils.py:17-18 (l1_loss) and :45-63 (_ssim, size_average=True) for [C,H,W] inputs.
    Returns (l1, ssim, dl1/dimg, dssim/dimg) with the analytic gradients autograd would produce."""
    x, y = img.astype(np.float64), gt.astype(np.float64)
    g = gaussian_window().astype(np.float64)
    w2 = np.outer(g.astype(np.float32), g.astype(np.float32)).astype(np.float32).astype(np.float64)  # :29 (fp32 mm)
    n = x.size
    mu1, mu2 = _conv(x, w2), _conv(y, w2)
    e11, e22, e12 = _conv(x * x, w2), _conv(y * y, w2), _conv(x * y, w2)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s11, s22, s12 = e11 - mu1_sq, e22 - mu2_sq, e12 - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    num1, num2 = 2 * mu12 + C1, 2 * s12 + C2
    den1, den2 = mu1_sq + mu2_sq + C1, s11 + s22 + C2
    smap = num1 * num2 / (den1 * den2)
    d_mu1 = 2 * mu2 * num2 / (den1 * den2) - smap * 2 * mu1 / den1
    d_s11 = -smap / den2
    d_s12 = 2 * num1 / (den1 * den2)
    A = d_mu1 - 2 * mu1 * d_s11 - mu2 * d_s12
    dssim = (_conv(A, w2) + 2 * x * _conv(d_s11, w2) + y * _conv(d_s12, w2)) / n
    dl1 = np.sign(x - y) / n
    return float(np.abs(x - y).mean()), float(smap.mean()), dl1, dssim


# ------------------------------------------------------------------------------ frustum
_PIECES = {(-1, 1): ((-1, -1, -1), (1, 1, 1))}
for _p in range(2):
    _PIECES[(_p, 2)] = ((0 if _p else -1, -1, -1), (1 if _p else 0, 1, 1))
for _p in range(4):
    _PIECES[(_p, 4)] = ((0 if _p & 1 else -1, 0 if _p & 2 else -1, -1), (1 if _p & 1 else 0, 1 if _p & 2 else 0, 1))


def in_frustum(proj, pts, faces=None, cube_len=None, piece_id=-1, n_piece=1):
    """finetune.py:33-48 / colmap.py:32-76 (+ centroids finetune.py:405 when faces is given)."""
    lo, hi = _PIECES[(piece_id, n_piece)] if piece_id >= 0 else _PIECES[(-1, 1)]
    proj = np.ascontiguousarray(proj, dtype=np.float32).reshape(16)
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    N = pts.shape[0] if faces is None else faces.shape[0]
    f = None if faces is None else np.ascontiguousarray(faces, dtype=np.int64)
    mask = np.zeros(N, dtype=np.uint8)
    lo_a, hi_a = np.array(lo, dtype=np.float32), np.array(hi, dtype=np.float32)
    O.lib().orc_in_frustum(C.c_int64(N), O._p(proj), C.c_float(cube_len or 0.0), C.c_int32(cube_len is not None),
                           O._p(lo_a), O._p(hi_a), O._p(pts), O._p(f), O._p(mask))
    return mask.astype(bool)


# ------------------------------------------------------------------------------ Adam
def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-15):
    """torch.optim.Adam single-tensor update (the optimiser the reference constructs at
    finetune.py:537 with eps=1e-15; no weight decay / amsgrad), fp32 state, scalars as torch rounds them."""
    f = np.float32
    p, g, m, v = (np.asarray(a, dtype=f) for a in (p, g, m, v))
    m = (m + f(1 - beta1) * (g - m)).astype(f)
    v = (v * f(beta2) + f(1 - beta2) * g * g).astype(f)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = (np.sqrt(v) / f(bc2 ** 0.5) + f(eps)).astype(f)
    p = (p - f(lr / bc1) * (m / denom)).astype(f)
    return p, m, v
