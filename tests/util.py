"""Shared helpers for the tests: scene builders and oracle parameter packing."""
import math
import os

import numpy as np
import torch

from dmgs_b200 import synthetic as S
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def cam_params(cam, P, bg, **kw):
    return O.make_params(P, cam.image_width, cam.image_height, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5),
                         bg, cam.world_view_transform.numpy(), cam.full_proj_transform.numpy(),
                         cam.camera_center.numpy(), **kw)


def small_scene(P=300, W=64, H=48, seed=3, scale=0.08, extent=1.0):
    cam = S.look_at_camera([2.5, 1.0, 1.2], W, H, fovx=0.9)
    cl = S.random_cloud(P, seed=seed, extent=extent, log_scale_mean=math.log(scale))
    return cam, cl


def cov6_from_scale_rot(scales, rotations):
    """utils/general_utils.py:78-110 semantics without the internal normalisation (unit quats given)."""
    from oracle import torch_oracle as TO
    R = TO.quat_to_rot(rotations.double())
    M = R @ torch.diag_embed(scales.double())
    Sg = M @ M.transpose(1, 2)
    return torch.stack([Sg[:, 0, 0], Sg[:, 0, 1], Sg[:, 0, 2], Sg[:, 1, 1], Sg[:, 1, 2], Sg[:, 2, 2]], 1).float()


def rel_err(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def grad_close(got, ref, rtol=1e-4, name=""):
    """Gradient parity at the north-star bar (1e-4 relative; accumulation order differs).

    |got-ref| <= rtol * max(scale, floor), where scale is the largest |ref| among the components
    of the same Gaussian (row) -- the components of one gradient vector are sums of the same
    large fp32 terms, so a small component next to a large one carries the large one's rounding --
    and floor = 1e-3 * RMS of the whole reference tensor (sums of mixed-sign fp32 terms cannot be
    relative-accurate near zero)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, f"{name}: shape {got.shape} vs {ref.shape}"
    g2, r2 = got.reshape(got.shape[0], -1), ref.reshape(ref.shape[0], -1)
    floor = 1e-3 * (np.sqrt(np.mean(r2 ** 2)) + 1e-30)
    scale = np.maximum(np.abs(r2).max(axis=1, keepdims=True), floor)
    err = np.abs(g2 - r2)
    bad = err > rtol * scale
    if bad.any():
        i, j = np.unravel_index(np.argmax(err / scale), err.shape)
        raise AssertionError(f"{name}: {bad.sum()} / {bad.size} elements off; worst row {i}: got {g2[i]} ref {r2[i]} "
                             f"(err {err[i, j]:.3e}, scale {scale[i, 0]:.3e})")


def grad_close_conditioned(got, ref, abs_sums, rtol=1e-4, ctol=1e-5, name=""):
    """Element-wise bar for gradients that are sums of many mixed-sign fp32 terms accumulated in an order that
    differs from the oracle's (and from run to run: atomics): |got - ref| <= rtol |ref| + ctol * (sum of the
    MAGNITUDES of the element's terms, the oracle's `abs9`), i.e. plain 1e-4 relative unless the sum cancels by
    more than 10x; no row scaling, no floor, no outliers."""
    got, ref, ab = (np.asarray(x, np.float64) for x in (got, ref, abs_sums))
    assert got.shape == ref.shape == ab.shape, f"{name}: shapes {got.shape} {ref.shape} {ab.shape}"
    err = np.abs(got - ref)
    bound = rtol * np.abs(ref) + ctol * ab + 1e-30
    worst = float((err / bound).max()) if err.size else 0.0
    assert worst <= 1.0, (f"{name}: worst error / bound = {worst:.3f} at "
                          f"{np.unravel_index(np.argmax(err / bound), err.shape)}")
