"""Fused L1 + SSIM image loss on the device (libdmgs_raster.so: dmgs_l1_ssim_forward/backward).

Host-side mirror of the reference's ``utils/loss_utils.py``: same names, arguments and results
  * ``l1_loss(network_output, gt)``                         <- loss_utils.py:17-18
  * ``ssim(img1, img2, window_size=11, size_average=True)`` <- loss_utils.py:35-63
plus the combination every trainer forms right after the render (train_geo_stage2.py:115-116):
  * ``l1_ssim_loss(image, gt, lambda_dssim)`` = (1 - lambda) * l1 + lambda * (1 - ssim), one forward
    kernel and one backward kernel instead of ~40 image-sized passes.
Inputs are [C,H,W] or [B,C,H,W] fp32 CUDA tensors; gradients flow to the first argument only (the
ground truth is data, as in the trainers).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from math import exp

import torch

from . import _lib as L


def gaussian_window(window_size: int = 11, sigma: float = 1.5):
    """The normalised 1-D window of loss_utils.py:23-25 (fp32 taps: torch.Tensor(list) / sum)."""
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    return g / g.sum()


_WIN = {}


def _window11():
    if "w" not in _WIN:
        g = gaussian_window(11, 1.5)
        _WIN["w"] = (C.c_float * 11)(*[float(v) for v in g])
    return _WIN["w"]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _planes(t: torch.Tensor):
    if t.dim() == 3:
        return 1, int(t.shape[0]), int(t.shape[1]), int(t.shape[2])
    if t.dim() == 4:
        return int(t.shape[0]), int(t.shape[1]), int(t.shape[2]), int(t.shape[3])
    raise ValueError(f"expected [C,H,W] or [B,C,H,W], got {tuple(t.shape)}")


class _L1SSIM(torch.autograd.Function):
    """Returns means [B*C, 2] = (mean |x-y|, mean ssim_map) per plane."""

    @staticmethod
    def forward(ctx, img, gt):
        if img.device.type != "cuda":
            raise RuntimeError("dmgs_b200 loss needs CUDA tensors; there is no CPU path")
        if img.shape != gt.shape:
            raise ValueError(f"shape mismatch {tuple(img.shape)} vs {tuple(gt.shape)}")
        B, Cn, H, W = _planes(img)
        x = img.detach().float().contiguous()
        y = gt.detach().float().contiguous()
        lib = L.lib()
        scratch = torch.empty(lib.dmgs_l1_ssim_scratch_bytes(B * Cn, H, W), dtype=torch.uint8, device=img.device)
        means = torch.empty(B * Cn, 2, dtype=torch.float32, device=img.device)
        L.check(lib.dmgs_l1_ssim_forward(B * Cn, H, W, _window11(), L.ptr(x), L.ptr(y), L.ptr(scratch), L.ptr(means),
                                         _stream()), "dmgs_l1_ssim_forward")
        ctx.save_for_backward(x, y, scratch)
        ctx.shape = (B * Cn, H, W, tuple(img.shape))
        return means

    @staticmethod
    def backward(ctx, g_means):
        x, y, scratch = ctx.saved_tensors
        planes, H, W, shape = ctx.shape
        up = (g_means.float() / float(H * W)).contiguous()
        grad = torch.empty(shape, dtype=torch.float32, device=x.device)
        L.check(L.lib().dmgs_l1_ssim_backward(planes, H, W, _window11(), L.ptr(x), L.ptr(y), L.ptr(scratch), L.ptr(up),
                                              L.ptr(grad), _stream()), "dmgs_l1_ssim_backward")
        return grad, None


def l1_ssim_means(img: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """[B*C, 2] per-plane (mean |img-gt|, mean ssim_map), differentiable w.r.t. img."""
    return _L1SSIM.apply(img, gt)


def l1_loss(network_output, gt):
    return l1_ssim_means(network_output, gt)[:, 0].mean()


def ssim(img1, img2, window_size=11, size_average=True):
    if window_size != 11:
        raise NotImplementedError("the fused kernel implements the 11-tap window the trainers use (loss_utils.py:35)")
    m = l1_ssim_means(img1, img2)[:, 1]
    if size_average:
        return m.mean()
    B, Cn, _, _ = _planes(img1)
    return m.view(B, Cn).mean(1)


def l1_ssim_loss(image, gt, lambda_dssim: float = 0.2):
    """(1 - lambda) * l1_loss + lambda * (1 - ssim)  (train_geo_stage2.py:115-116)."""
    m = l1_ssim_means(image, gt).mean(0)
    return (1.0 - lambda_dssim) * m[0] + lambda_dssim * (1.0 - m[1])


_UP = {}


def l1_ssim_loss_and_grad(image, gt, lambda_dssim: float = 0.2, need_loss: bool = True):
    """(loss, dloss/dimage) of ``l1_ssim_loss`` without the autograd round trip: three kernel launches
    (forward, the fixed-order reduction, backward).  This is the ``image_grad`` callback of the
    view-batched training step (multiview.accumulate_view).  loss is None when need_loss is False."""
    if image.device.type != "cuda":
        raise RuntimeError("dmgs_b200 loss needs CUDA tensors; there is no CPU path")
    B, Cn, H, W = _planes(image)
    planes = B * Cn
    x, y = image.detach().float().contiguous(), gt.detach().float().contiguous()
    lib, dev = L.lib(), image.device
    scratch = torch.empty(lib.dmgs_l1_ssim_scratch_bytes(planes, H, W), dtype=torch.uint8, device=dev)
    means = torch.empty(planes, 2, dtype=torch.float32, device=dev)
    key = (dev.index, planes, H, W, float(lambda_dssim))
    up = _UP.get(key)
    if up is None:  # d loss / d (per-plane means), divided by the pixel count: constants of the shape
        up = torch.tensor([(1.0 - lambda_dssim) / planes / (H * W), -lambda_dssim / planes / (H * W)],
                          dtype=torch.float32).repeat(planes, 1).to(dev)
        _UP[key] = up
    grad = torch.empty_like(x)
    st = _stream()
    L.check(lib.dmgs_l1_ssim_forward(planes, H, W, _window11(), L.ptr(x), L.ptr(y), L.ptr(scratch), L.ptr(means), st),
            "dmgs_l1_ssim_forward")
    L.check(lib.dmgs_l1_ssim_backward(planes, H, W, _window11(), L.ptr(x), L.ptr(y), L.ptr(scratch), L.ptr(up), L.ptr(grad),
                                      st), "dmgs_l1_ssim_backward")
    loss = None
    if need_loss:
        m = means.mean(0)
        loss = (1.0 - lambda_dssim) * m[0] + lambda_dssim * (1.0 - m[1])
    return loss, grad
