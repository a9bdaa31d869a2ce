"""Full-size GPU tests (BASELINE.json shapes: 1 M Gaussians at 800x800, 1245x825 and 1920x1080; the C2 mesh-bound
stage-2 step as a chain).  Two kinds:
  (A) against the CPU oracle on the same seeded inputs (one frame each, a few seconds of CPU): forward bit-exact
      (radii, tiles, depth / xy / conic / colour bits, R, 64-bit keys, values, ranges, n_contrib, final T), image
      <= 1e-5, every gradient within the 1e-4 bar (tests/util.py:grad_close);
  (B) size-independent properties:
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


# ------------------------------------------------------------------------ (A) against the CPU oracle
def _oracle_parity(name, backward):
    import numpy as np
    from oracle import oracle as O
    from test_gpu_parity import assert_forward_parity, run_both
    from util import grad_close
    from dmgs_b200.rasterizer import rasterize_backward
    P, W, H, kind, extent, lsm = SHAPES[name]
    cl = S.random_cloud(P, seed=0, extent=extent, log_scale_mean=lsm)
    cam = S.nerf_synthetic_camera(1, W, H) if kind == "nerf" else S.bicycle_camera(1, W, H)
    pr, npin, ref, color, radii, st, arrays = run_both(cam, cl, "sh_scale_rot", bg=(0.1, 0.2, 0.3))
    assert ref["bins"]["R"] > P
    assert_forward_parity(ref, color, radii, st, arrays)
    if not backward:
        return
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(5))
    bw = O.render_backward(pr, ref, dL.numpy(), cl["means3D"].numpy(), scales=npin["scales"], rotations=npin["rotations"],
                           shs=npin["shs"], abs_sums=True)
    d = {k: torch.tensor(v).cuda() for k, v in npin.items()}
    g = rasterize_backward(st, dL.cuda(), cl["means3D"].cuda(), d["shs"], d["scales"], d["rotations"], None, False)
    torch.cuda.synchronize()
    got = dict(zip(["means3D", "means2D", "shs", "col", "opacity", "scales", "rotations", "cov"],
                   [None if t is None else t.cpu().numpy() for t in g]))
    # (1) the blend backward (K7), element by element with NO outliers allowed.  Every one of its outputs is a sum
    #     over up to thousands of (pixel, Gaussian) terms of mixed sign, accumulated in fp32 in an order that differs
    #     from the oracle's (and from run to run: atomics) -- so the bar is 1e-4 relative to the gradient plus 1e-5 of
    #     the sum of the MAGNITUDES of its terms (oracle: abs9), i.e. plain 1e-4 relative unless the sum cancels >10x
    gb = st.ws.scratch[:P * 48].view(torch.float32).view(P, 12).cpu().numpy().astype(np.float64)
    ref9 = np.concatenate([bw["dL_dmean2D"], bw["dL_dconic"], bw["dL_dopacity"][:, None], bw["dL_dcolor"]], 1).astype(np.float64)
    err9 = np.abs(gb[:, :9] - ref9)
    bound9 = 1e-4 * np.abs(ref9) + 1e-5 * bw["abs9"].astype(np.float64) + 1e-30
    worst = float((err9 / bound9).max())
    assert worst <= 1.0, f"blend backward: worst error / bound = {worst:.3f} at {np.unravel_index(np.argmax(err9 / bound9), err9.shape)}"
    # (2) the per-Gaussian backward (K8 + K9) on ITS inputs: the oracle's per-Gaussian backward fed with the blend
    #     gradients the GPU produced must give the GPU's final gradients (no accumulation-order freedom left)
    gpu_blend = {"dL_dmean2D": gb[:, 0:2].astype(np.float32), "dL_dconic": gb[:, 2:5].astype(np.float32),
                 "dL_dopacity": gb[:, 5].astype(np.float32), "dL_dcolor": gb[:, 6:9].astype(np.float32)}
    pg = O.preprocess_backward(pr, ref["geom"], gpu_blend, cl["means3D"].numpy(), npin["scales"], npin["rotations"], npin["shs"])
    grad_close(got["means3D"], pg["dL_dmeans3D"], name="means3D | gpu blend grads")
    grad_close(got["shs"], pg["dL_dshs"], name="shs | gpu blend grads")
    grad_close(got["scales"], pg["dL_dscales"], name="scales | gpu blend grads")
    grad_close(got["rotations"], pg["dL_drotations"], name="rotations | gpu blend grads")
    assert np.array_equal(got["means2D"][:, :2], gpu_blend["dL_dmean2D"]) and (got["means2D"][:, 2] == 0).all()
    assert np.array_equal(got["opacity"][:, 0], gpu_blend["dL_dopacity"])
    # (3) end to end against the oracle: norm-wise 1e-5 (ten times inside the 1e-4 bar), and element-wise at the
    #     row-scaled 1e-4 bar of tests/util.py for all but a 1e-4 fraction of the elements (the ill-conditioned sums of (1))
    pairs = [("means3D", bw["dL_dmeans3D"]), ("means2D", np.concatenate([bw["dL_dmean2D"], np.zeros((P, 1), np.float32)], 1)),
             ("opacity", bw["dL_dopacity"][:, None]), ("shs", bw["dL_dshs"]), ("scales", bw["dL_dscales"]),
             ("rotations", bw["dL_drotations"])]
    for name_, r in pairs:
        a_, r_ = got[name_].astype(np.float64).reshape(P, -1), r.astype(np.float64).reshape(P, -1)
        assert np.isfinite(a_).all(), name_
        rel = np.linalg.norm(a_ - r_) / np.linalg.norm(r_)
        assert rel <= 1e-5, f"{name_}: norm-wise relative error {rel:.2e}"
        floor = 1e-3 * np.sqrt(np.mean(r_ ** 2))
        scale = np.maximum(np.abs(r_).max(axis=1, keepdims=True), floor)
        frac = float((np.abs(a_ - r_) > 1e-4 * scale).mean())
        assert frac <= 1e-4, f"{name_}: {frac:.2e} of the elements miss the row-scaled 1e-4 bar"


@pytest.mark.parametrize("name", ["h0", "c3"])
def test_full_size_forward_and_backward_vs_oracle(name):
    """BASELINE.json's headline shape (H0) and configs[2] (C3): forward bit-exact, backward within 1e-4."""
    _oracle_parity(name, backward=True)


def test_full_size_forward_vs_oracle_1080p():
    """configs[4]'s largest image (1 M Gaussians at 1920x1080, 8160 tiles): forward bit-exact."""
    _oracle_parity("c5_1080p", backward=False)


_C2 = {}


def _c2_oracle_chain():
    """The C2 inputs and the oracle's chain on them (computed once per session): binding -> render (sigmoid SH in
    preprocess, cov3D_precomp, opacity 0.9999) -> 0.8 L1 + 0.2 (1 - SSIM) -> backward -> dverts, dg, dfeatures."""
    if _C2:
        return _C2
    import numpy as np
    from oracle import next_rows as NR, oracle as O
    from util import cam_params
    m = S.mesh_bound_inputs(50_000, k=6, seed=1)
    W = H = 800
    cam = S.nerf_synthetic_camera(2, W, H)
    bg = (1.0, 1.0, 1.0)
    thin_z = m["spatial_lr_scale"] * 1e-6
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(11))
    sf = torch.tensor([m["scale_factor"]], device="cuda")
    g = float((torch.tanh(sf) * 2.0).item())  # the fp32 value the kernels read
    ob = O.bind_forward(m["verts"].numpy(), m["faces"].numpy(), m["bc"].numpy(), m["rad_base"], thin_z, g, True)
    P = ob["xyz"].shape[0]
    pr = cam_params(cam, P, np.array(bg, np.float32), sh_layout=1, sh_act=1, sh_degree=3)
    ref = O.render_forward(pr, ob["xyz"], m["opacities"].numpy(), cov3D_precomp=ob["cov6"], shs=m["features"].numpy())
    l1, ssim, dl1, dssim = NR.l1_ssim(ref["img"]["color"], gt.numpy())
    dL = (0.8 * dl1 - 0.2 * dssim).astype(np.float32)
    bw = O.render_backward(pr, ref, dL, ob["xyz"], shs=m["features"].numpy())
    ov = O.bind_backward(m["verts"].numpy(), m["faces"].numpy(), m["bc"].numpy(), m["rad_base"], thin_z, g,
                         bw["dL_dmeans3D"], bw["dL_dcov3D"], True)
    _C2.update(m=m, cam=cam, bg=bg, thin_z=thin_z, gt=gt, g=g, ob=ob, P=P, pr=pr, ref=ref, bw=bw, ov=ov,
               loss=0.8 * l1 + 0.2 * (1.0 - ssim),
               dsf=ov["dg"] * 2.0 * (1.0 - math.tanh(m["scale_factor"]) ** 2))
    return _C2


def _c2_check(c, loss, img, radii, state, v, sf, feats):
    import numpy as np
    from test_gpu_parity import assert_forward_parity
    from util import grad_close
    from gpu_util import state_numpy
    assert_forward_parity(c["ref"], img.detach(), radii, state, state_numpy(state))
    assert abs(loss.item() - c["loss"]) <= 1e-5
    P = c["P"]
    grad_close(feats.grad.cpu().numpy().reshape(P, -1), c["bw"]["dL_dshs"].reshape(P, -1), name="features")
    grad_close(v.grad.cpu().numpy(), c["ov"]["dverts"], rtol=2e-4, name="dverts")
    assert abs(sf.grad.item() - c["dsf"]) <= 2e-4 * abs(c["dsf"]) + 1e-12
    assert np.isfinite(v.grad.cpu().numpy()).all()


def test_c2_stage2_chain_vs_oracle():
    """BASELINE.json configs[1] as a CHAIN against the oracle chain: mesh (F ~ 82 k, k = 6 -> P ~ 490 k) -> binding
    -> sigmoid-SH colour inside preprocess + cov3D_precomp render (opacity 0.9999: alpha saturates at 0.99) ->
    0.8 L1 + 0.2 (1 - SSIM) -> backward -> dL/dverts, dL/dscale_factor, dL/dfeatures."""
    import numpy as np
    from gpu_util import settings_for
    from dmgs_b200 import GaussianRasterizer
    from dmgs_b200.binding import bind_faces
    from dmgs_b200.loss_utils import l1_ssim_loss
    c = _c2_oracle_chain()
    m = c["m"]
    v = m["verts"].cuda().requires_grad_()
    sf = torch.tensor([m["scale_factor"]], device="cuda", requires_grad=True)
    feats = m["features"].cuda().requires_grad_()
    xyz, cov6 = bind_faces(v, m["faces"].cuda(), m["bc"].cuda(), m["rad_base"], c["thin_z"], sf, max_scale=2.0)
    assert np.array_equal(xyz.detach().cpu().numpy().view(np.uint32), c["ob"]["xyz"].view(np.uint32))
    assert np.array_equal(cov6.detach().cpu().numpy().view(np.uint32), c["ob"]["cov6"].view(np.uint32))
    ras = GaussianRasterizer(settings_for(c["cam"], c["bg"]), sh_activation="sigmoid", sh_layout="P3M")
    img, radii = ras(means3D=xyz, means2D=torch.zeros_like(xyz), shs=feats, colors_precomp=None,
                     opacities=m["opacities"].cuda(), scales=None, rotations=None, cov3D_precomp=cov6)
    loss = l1_ssim_loss(img, c["gt"].cuda(), lambda_dssim=0.2)
    loss.backward()
    torch.cuda.synchronize()
    _c2_check(c, loss, img, radii, ras.last, v, sf, feats)


def test_c2_fused_bind_preprocess_vs_oracle():
    """The same chain with the binding fused INTO the per-Gaussian kernels (dmgs_preprocess_forward_bound /
    dmgs_preprocess_backward_bound via binding.rasterize_mesh): no xyz / cov3D_precomp / dL/dxyz / dL/dcov tensors.
    Forward bit-exact against the oracle chain, gradients (the reference's truncated binding gradient) within the bar."""
    import numpy as np
    from gpu_util import settings_for
    from dmgs_b200.binding import MeshRasterizer, rasterize_mesh
    from dmgs_b200.loss_utils import l1_ssim_loss
    c = _c2_oracle_chain()
    m = c["m"]
    v = m["verts"].cuda().requires_grad_()
    sf = torch.tensor([m["scale_factor"]], device="cuda", requires_grad=True)
    feats = m["features"].cuda().requires_grad_()
    holder = MeshRasterizer()
    m2d = torch.zeros(c["P"], 3, device="cuda", requires_grad=True)
    img, radii, xyz = rasterize_mesh(settings_for(c["cam"], c["bg"]), v, m["faces"].cuda(), m["bc"].cuda(), m["rad_base"],
                                     c["thin_z"], m["opacities"].cuda(), features=feats, scale_factor=sf, max_scale=2.0,
                                     means2D=m2d, return_xyz=True, holder=holder)
    assert np.array_equal(xyz.cpu().numpy().view(np.uint32), c["ob"]["xyz"].view(np.uint32))
    loss = l1_ssim_loss(img, c["gt"].cuda(), lambda_dssim=0.2)
    loss.backward()
    torch.cuda.synchronize()
    _c2_check(c, loss, img, radii, holder.last, v, sf, feats)
    assert m2d.grad.shape == (c["P"], 3) and torch.all(m2d.grad[:, 2] == 0) and m2d.grad.abs().sum() > 0
    # the COLMAP variant (no scale factor, mlp_flex_colmap.py:500-504) and precomputed colours run too
    col = torch.rand(c["P"], 3, device="cuda", requires_grad=True)
    v2 = m["verts"].cuda().requires_grad_()
    img2, radii2 = rasterize_mesh(settings_for(c["cam"], c["bg"]), v2, m["faces"].cuda(), m["bc"].cuda(), m["rad_base"],
                                  c["thin_z"], m["opacities"].cuda(), colors_precomp=col)
    img2.square().mean().backward()
    assert torch.isfinite(v2.grad).all() and v2.grad.abs().sum() > 0 and col.grad.abs().sum() > 0


def test_arith_divergence_script_small():
    """scripts/arith_divergence.py (product contract vs upstream-grouping/nvcc-default arithmetic) runs and the two
    arithmetics agree on almost every decision of a small frame (sanity of the measuring tool, not a parity bar)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("arith_divergence", os.path.join(root, "scripts", "arith_divergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.compare("small")
    assert r["visible_ours"] > 1000
    assert r["culling_decision_differs"] <= 2 and r["radius_differs"] <= r["visible_ours"] // 100
    assert r["image_max_abs_diff"] < 2e-2


def test_full_size_colour_linearity_and_adjoint_h0():
    """H0 shape, precomputed colours: the image is linear in the colours and affine in the background, the weights
    (final T, contributor counts) do not depend on the colours, and the backward's colour gradient is the adjoint of
    that linear map: sum_i <dL/dc_i, c_i> = <dL/dimage, image - final_T * bg> (the same identities the CPU oracle is held
    to in tests/test_oracle.py)."""
    from dmgs_b200.rasterizer import rasterize_backward, rasterize_forward
    d, rs, P, W, H = _scene("h0")
    gen = torch.Generator().manual_seed(3)
    c1 = torch.rand(P, 3, generator=gen).cuda()
    c2 = torch.rand(P, 3, generator=gen).cuda()
    dL = torch.randn(3, H, W, generator=gen).cuda()
    bg = rs.bg.view(3, 1, 1)

    def render(col):
        color, radii, st = rasterize_forward(rs, d["means3D"], d["opacities"], None, col.contiguous(), d["scales"],
                                             d["rotations"], None)
        im = st.image_arrays()
        return color, im["final_T"].clone(), im["n_contrib"].clone(), st

    i1, T1, n1, st1 = render(c1)
    g = rasterize_backward(st1, dL, d["means3D"], None, d["scales"], d["rotations"], None, True)
    g_col = g[3]
    i2, T2, n2, _ = render(c2)
    i12, _, _, _ = render(0.25 * c1 + 1.5 * c2)
    torch.cuda.synchronize()
    assert torch.equal(T1, T2) and torch.equal(n1, n2)
    lin = (i12 - bg * T1) - (0.25 * (i1 - bg * T1) + 1.5 * (i2 - bg * T1))
    assert float(lin.abs().max()) < 1e-4, float(lin.abs().max())
    lhs = float((g_col.double() * c1.double()).sum())
    rhs = float((dL.double() * (i1 - bg * T1).double()).sum())
    scale = float((dL.abs().double() * (i1 - bg * T1).abs().double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * scale, (lhs, rhs, scale)
    lhs2 = float((g_col.double() * c2.double()).sum())
    rhs2 = float((dL.double() * (i2 - bg * T1).double()).sum())
    assert abs(lhs2 - rhs2) <= 1e-4 * scale, (lhs2, rhs2, scale)
