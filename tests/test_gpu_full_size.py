"""Full-size GPU tests (BASELINE.json shapes: 1 M Gaussians at 800x800, 1245x825 and 1920x1080) through
size-independent properties -- the CPU oracle takes too long at these sizes:
  * the rebuilt 64-bit (tile | depth) keys are sorted and every key's Gaussian really touches its tile;
  * the tile ranges partition [0, R) exactly along the key's tile ids; R = sum(tiles_touched);
  * contributor counts never exceed the tile's list length; final T in (0, 1]; image finite and bounded;
  * forward is bit-reproducible run to run;
  * the backward is linear in dL/dimage (the adjoint of a fixed forward): grad(a*d1 + b*d2) = a*grad(d1) +
    b*grad(d2) within the 1e-4 gradient bar, and a directional finite difference of the forward along the SH
    coefficients agrees with <grad, direction>."""
import math

import pytest
import torch

from dmgs_b200 import synthetic as S

pytestmark = pytest.mark.gpu

SHAPES = {
    "h0": (1_000_000, 800, 800, "nerf", 1.3, math.log(0.01)),
    "c3": (1_000_000, 1245, 825, "bicycle", 3.0, math.log(0.008)),
    "c5_1080p": (1_000_000, 1920, 1080, "nerf", 1.3, math.log(0.01)),
}


def _scene(name):
    from gpu_util import settings_for
    P, W, H, kind, extent, lsm = SHAPES[name]
    cl = S.random_cloud(P, seed=0, extent=extent, log_scale_mean=lsm)
    cam = S.nerf_synthetic_camera(1, W, H) if kind == "nerf" else S.bicycle_camera(1, W, H)
    return {k: v.cuda() for k, v in cl.items()}, settings_for(cam, (0.1, 0.2, 0.3)), P, W, H


@pytest.mark.parametrize("name", list(SHAPES))
def test_full_size_forward_invariants(name):
    from dmgs_b200.rasterizer import rasterize_forward
    d, rs, P, W, H = _scene(name)
    args = (rs, d["means3D"], d["opacities"], d["shs"], None, d["scales"], d["rotations"], None)
    color, radii, st = rasterize_forward(*args)
    torch.cuda.synchronize()
    R = st.num_rendered
    g, b, im = st.geom_arrays(), st.binning_arrays(), st.image_arrays()
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    tiles_touched = g["tiles_touched"].long()
    assert R == int(tiles_touched.sum()) and R > 0
    assert torch.equal(radii > 0, tiles_touched > 0) or int(((radii > 0) != (tiles_touched > 0)).sum()) == int(
        ((radii > 0) & (tiles_touched == 0)).sum())  # a visible radius may still cover no tile centre-wise
    keys = st.sorted_keys()                                       # int64 view of the u64 keys
    assert bool((keys[1:] >= keys[:-1]).all()), "keys not sorted"  # tile ids < 2^31: signed compare is safe
    key_tile = (keys >> 32).int()
    gidx = b["gidx"].long()
    # every instance's Gaussian covers its tile, and carries that Gaussian's depth bits
    rect = g["rect"].int()[gidx]                                  # x0, x1, y0, y1 as stored (u16 pairs)
    tx, ty = key_tile % gx, key_tile // gx
    inside = (tx >= rect[:, 0]) & (tx < rect[:, 1]) & (ty >= rect[:, 2]) & (ty < rect[:, 3])
    assert bool(inside.all()), "an instance lies outside its Gaussian's tile rectangle"
    depth_bits = g["depths"].view(torch.int32)[gidx].long() & 0xFFFFFFFF
    assert torch.equal(keys & 0xFFFFFFFF, depth_bits)
    # ranges: [start, end) of each tile id in the sorted keys; empty tiles keep (0, 0)
    counts = torch.bincount(key_tile.long(), minlength=T)
    ends = torch.cumsum(counts, 0)
    starts = ends - counts
    rng = b["ranges"].long()
    nonempty = counts > 0
    assert torch.equal(rng[nonempty, 0], starts[nonempty]) and torch.equal(rng[nonempty, 1], ends[nonempty])
    assert bool((rng[~nonempty] == 0).all())
    # image-side invariants
    n_contrib = im["n_contrib"].long().view(gy * 0 + H, W)
    tile_of_pixel = (torch.arange(H, device="cuda")[:, None] // 16) * gx + (torch.arange(W, device="cuda")[None, :] // 16)
    assert bool((n_contrib <= counts[tile_of_pixel]).all())
    fT = im["final_T"]
    assert bool(((fT > 0) & (fT <= 1)).all())
    assert bool(torch.isfinite(color).all()) and float(color.min()) >= 0.0
    # bit-reproducible
    color2, radii2, st2 = rasterize_forward(*args)
    assert torch.equal(color, color2) and torch.equal(radii, radii2)
    assert torch.equal(st2.binning_arrays()["gidx"], b["gidx"]) and torch.equal(st2.image_arrays()["n_contrib"], im["n_contrib"])


def test_full_size_backward_linearity_and_directional_derivative():
    from dmgs_b200.rasterizer import rasterize_backward, rasterize_forward
    d, rs, P, W, H = _scene("h0")
    color, radii, st = rasterize_forward(rs, d["means3D"], d["opacities"], d["shs"], None, d["scales"], d["rotations"], None)
    gen = torch.Generator().manual_seed(3)
    d1 = torch.randn(3, H, W, generator=gen).cuda()
    d2 = torch.randn(3, H, W, generator=gen).cuda()
    bw = lambda dl: rasterize_backward(st, dl, d["means3D"], d["shs"], d["scales"], d["rotations"], None, False)
    g1, g2, g12 = bw(d1), bw(d2), bw(0.75 * d1 - 1.5 * d2)
    for a, b, c, name in zip(g1, g2, g12, ["means3D", "means2D", "shs", "col", "opacities", "scales", "rotations", "cov"]):
        if a is None:
            continue
        lin = 0.75 * a - 1.5 * b
        num = torch.linalg.norm((c - lin).double())
        den = torch.linalg.norm(lin.double())
        assert float(num) <= 1e-4 * float(den), f"{name}: backward not linear in dL/dimage ({float(num / den):.2e})"
    # directional derivative of sum(image * d1) along a perturbation of the SH coefficients: the image is
    # continuous and piecewise linear in them (the alpha / transmittance thresholds, which make the image
    # discontinuous in opacity and geometry, do not depend on colour)
    t = 1e-2
    dsh = g1[2] * (0.05 / float(g1[2].abs().max()))  # along the gradient: a well-conditioned inner product
    f = lambda s: float((rasterize_forward(rs, d["means3D"], d["opacities"], d["shs"] + s * dsh, None, d["scales"],
                                            d["rotations"], None)[0].double() * d1.double()).sum())
    num = (f(t) - f(-t)) / (2 * t)
    ana = float((g1[2].double() * dsh.double()).sum())
    assert abs(num - ana) <= 1e-2 * abs(ana) + 1e-2, (num, ana)
