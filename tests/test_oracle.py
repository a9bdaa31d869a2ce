"""CPU tests: the oracle against the reference's golden vectors and against fp64 autograd."""
import math

import numpy as np
import pytest
import torch

from dmgs_b200 import synthetic as S
from oracle import oracle as O
from oracle import torch_oracle as TO
from util import cam_params, cov6_from_scale_rot, golden, rel_err, small_scene


def test_exp_contract_accuracy():
    x = np.concatenate([np.linspace(-90, 5, 200001), [-1e30, 0.0, -0.0, 100.0]]).astype(np.float32)
    y = O.exp_array(x)
    ref = np.exp(np.clip(x.astype(np.float64), -80, 80))
    assert np.max(np.abs(y - ref) / ref) < 3e-7
    assert O.exp_array(np.array([0.0], np.float32))[0] == 1.0


def test_barycentric_golden():
    g = golden("barycentric.npz")
    for k in (1, 3, 6):
        bc, rad = S.barycentric_layout(k)
        assert np.array_equal(bc.numpy(), g[f"bc{k}"])
        assert rad == float(g[f"rad{k}"])
    assert math.isclose(S.barycentric_layout(3)[1], 0.18301270189221933)
    assert math.isclose(S.barycentric_layout(6)[1], 0.13397459621556135)


def test_camera_golden():
    g = golden("camera.npz")
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = g["R"].T
    Rt[:3, 3] = g["T"]
    Rt[3, 3] = 1
    c2w = np.linalg.inv(Rt)
    c2w[:3, 1:3] *= -1  # camera_from_c2w takes the OpenGL convention
    cam = S.camera_from_c2w(c2w, 100, 80, float(g["fovx"]), float(g["fovy"]))
    assert np.allclose(cam.world_view_transform.numpy(), g["world_view"], atol=1e-6)
    assert np.allclose(cam.full_proj_transform.numpy(), g["full_proj"], atol=1e-5)
    assert np.allclose(cam.camera_center.numpy(), g["center"], atol=1e-5)


@pytest.mark.parametrize("k", [1, 3, 6])
@pytest.mark.parametrize("adaptive", [True, False])
def test_binding_stage2_golden(k, adaptive):
    g = golden("binding_stage2.npz")
    tag = f"k{k}_{'adp' if adaptive else 'iso'}"
    bc, rad = S.barycentric_layout(k)
    sf = float(g[f"{tag}_scale_factor"][0])
    gscale = math.tanh(sf) * 2
    out = O.bind_forward(g[f"{tag}_verts"], g[f"{tag}_faces"], bc.numpy(), rad, 4.43 * 1e-6, gscale, adaptive)
    assert np.allclose(out["xyz"], g[f"{tag}_xyz"], rtol=1e-5, atol=1e-6)
    assert np.allclose(out["rot_t2w"], g[f"{tag}_rot_t2w"], atol=2e-6)
    assert np.allclose(out["cov3D_L"], g[f"{tag}_cov3D_L"], rtol=2e-5, atol=1e-9)
    assert rel_err(out["cov6"], g[f"{tag}_cov6"]) < 2e-6
    bw = O.bind_backward(g[f"{tag}_verts"], g[f"{tag}_faces"], bc.numpy(), rad, 4.43 * 1e-6, gscale,
                         g[f"{tag}_gxyz"], g[f"{tag}_gcov"], adaptive)
    assert rel_err(bw["dverts"], g[f"{tag}_dverts"]) < 2e-5
    dsf = bw["dg"] * 2 * (1 - math.tanh(sf) ** 2)
    assert abs(dsf - float(g[f"{tag}_dscale_factor"][0])) <= 2e-5 * abs(float(g[f"{tag}_dscale_factor"][0]))


def test_affine_known_answers():
    # geo/affine_verify.ipynb cells 0-1 (example B), 4-dp values from the notebook
    verts = np.array([[0, 0, 0], [0.0038, 0, 0], [0.0011, 0.0035, 0]], np.float32)
    bc, rad = S.barycentric_layout(6)
    out = O.bind_forward(verts, np.array([[0, 1, 2]]), bc.numpy(), rad, 4.43e-6, 1.0, True)
    L = out["cov3D_L"][0]
    assert np.allclose(np.round(L[:2, :2], 4), [[0.0005, -0.0001], [0.0, 0.0005]])
    assert np.allclose(L, golden("binding_stage2.npz")["affineB_cov3D_L"][0], rtol=1e-5, atol=1e-10)
    # geo/affine_proto.ipynb cells 0-1 (example A): M_2d for tri (0,0),(1,0),(0.9,0.6)
    verts = np.array([[0, 0, 0], [1, 0, 0], [0.9, 0.6, 0]], np.float32)
    out = O.bind_forward(verts, np.array([[0, 1, 2]]), bc.numpy(), rad, 0.0, 1.0, True)
    M = out["cov3D_L"][0][:2, :2] / rad  # l = 1
    h = math.sqrt(3) / 2
    assert np.allclose(M, [[1, (0.9 - 0.5) / h], [0, 0.6 / h]], rtol=1e-5)


def _sh_scene(P):
    # camera far enough that every golden point is visible
    g = golden("eval_sh.npz")
    cam = S.look_at_camera([0.0, -9.0, 0.0], 64, 64, fovx=1.2)
    return g, cam


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
@pytest.mark.parametrize("act", [0, 1])
def test_eval_sh_golden(deg, act):
    g, cam = _sh_scene(64)
    P = g["xyz"].shape[0]
    pr = cam_params(cam, P, [0, 0, 0], sh_degree=deg, sh_layout=1, sh_act=act)
    pr.campos[:] = [float(v) for v in g["campos"]]  # colour uses campos only
    geom = O.preprocess(pr, g["xyz"], np.full((P, 1), 0.5, np.float32), scales=np.full((P, 3), 0.01, np.float32),
                        rotations=np.tile(np.array([[1, 0, 0, 0]], np.float32), (P, 1)), shs=g["features"])
    vis = geom["radii"] > 0
    assert vis.sum() > P // 2
    want = g[f"{'clamp' if act == 0 else 'sigmoid'}{deg}"]
    assert np.allclose(geom["rgb"][vis], want[vis], atol=2e-6)
    # same numbers through the [P,M,3] rasteriser layout
    pr2 = cam_params(cam, P, [0, 0, 0], sh_degree=deg, sh_layout=0, sh_act=act)
    pr2.campos[:] = [float(v) for v in g["campos"]]
    geom2 = O.preprocess(pr2, g["xyz"], np.full((P, 1), 0.5, np.float32), scales=np.full((P, 3), 0.01, np.float32),
                         rotations=np.tile(np.array([[1, 0, 0, 0]], np.float32), (P, 1)),
                         shs=np.ascontiguousarray(g["features"].transpose(0, 2, 1)))
    assert np.array_equal(geom["rgb"], geom2["rgb"])
    # and the fp64 restatement
    t = TO.eval_sh_colors(deg, torch.tensor(g["features"]).double().transpose(1, 2), torch.tensor(g["xyz"]).double(),
                          torch.tensor(g["campos"]).double(), act)
    assert np.allclose(t.numpy(), want, atol=1e-6)
    assert np.all(g["zero"] == 0)


def test_cov_layout_golden():
    g = golden("cov_layout.npz")
    q = g["q"] / np.linalg.norm(g["q"], axis=1, keepdims=True)  # callers normalise (gaussian_model.py:101)
    P = q.shape[0]
    cam = S.look_at_camera([0.0, -6.0, 0.0], 64, 64, fovx=1.2)
    pr = cam_params(cam, P, [0, 0, 0])
    xyz = np.zeros((P, 3), np.float32)
    xyz[:, 0] = np.linspace(-1, 1, P)
    geom = O.preprocess(pr, xyz, np.full((P, 1), 0.5, np.float32), scales=g["s"], rotations=q.astype(np.float32),
                        colors_precomp=np.zeros((P, 3), np.float32))
    assert (geom["radii"] > 0).all()
    assert rel_err(geom["cov3D"], g["cov6"]) < 2e-6


MODES = ["sh_scale_rot", "precomp", "sigmoid_features"]


def _mode_inputs(mode, cl, P):
    if mode == "sh_scale_rot":
        return dict(scales=cl["scales"], rotations=cl["rotations"], shs=cl["shs"]), {}
    if mode == "precomp":
        col = torch.rand(P, 3, generator=torch.Generator().manual_seed(9))
        return dict(cov3D_precomp=cov6_from_scale_rot(cl["scales"], cl["rotations"]), colors_precomp=col), {}
    feats = cl["shs"].transpose(1, 2).contiguous()
    return dict(scales=cl["scales"], rotations=cl["rotations"], shs=feats), dict(sh_layout=1, sh_act=1, sh_degree=2)


@pytest.mark.parametrize("mode", MODES)
def test_c_oracle_matches_fp64_autograd(mode):
    cam, cl = small_scene()
    P, W, H = cl["means3D"].shape[0], cam.image_width, cam.image_height
    inp, pk = _mode_inputs(mode, cl, P)
    bg = np.array([0.2, 0.5, 0.7], np.float32)
    pr = cam_params(cam, P, bg, **pk)
    npin = {k: v.numpy() for k, v in inp.items()}
    fwd = O.render_forward(pr, cl["means3D"].numpy(), cl["opacities"].numpy(), **npin)
    assert fwd["bins"]["R"] > 500
    m3 = cl["means3D"].double().requires_grad_()
    op = cl["opacities"].double().requires_grad_()
    tin = {k: v.double().requires_grad_() for k, v in inp.items()}
    img, inter = TO.render(W, H, pr.tanfovx, pr.tanfovy, torch.tensor(bg), cam.world_view_transform,
                           cam.full_proj_transform, cam.camera_center, m3, op, **tin, **pk)
    assert np.array_equal(fwd["geom"]["radii"], inter["radii"].numpy())
    assert np.abs(fwd["img"]["color"] - img.detach().numpy()).max() < 5e-6
    assert np.abs(fwd["img"]["final_T"] - inter["final_T"].detach().numpy()).max() < 5e-6
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(5))
    (img * dL.double()).sum().backward()
    bw = O.render_backward(pr, fwd, dL.numpy(), cl["means3D"].numpy(), scales=npin.get("scales"),
                           rotations=npin.get("rotations"), shs=npin.get("shs"),
                           precomp_color="colors_precomp" in npin)
    assert rel_err(bw["dL_dmeans3D"], m3.grad) < 2e-5
    assert rel_err(bw["dL_dopacity"], op.grad) < 2e-5
    if "scales" in tin:
        assert rel_err(bw["dL_dscales"], tin["scales"].grad) < 2e-5
        assert rel_err(bw["dL_drotations"], tin["rotations"].grad) < 2e-5
    if "shs" in tin:
        assert rel_err(bw["dL_dshs"], tin["shs"].grad) < 2e-5
    if "cov3D_precomp" in tin:
        assert rel_err(bw["dL_dcov3D"], tin["cov3D_precomp"].grad) < 2e-5
        assert rel_err(bw["dL_dcolors_precomp"], tin["colors_precomp"].grad) < 2e-5


def test_binning_properties():
    cam, cl = small_scene(P=2000, W=160, H=96, scale=0.05)
    P = 2000
    pr = cam_params(cam, P, [0, 0, 0])
    geom = O.preprocess(pr, cl["means3D"].numpy(), cl["opacities"].numpy(), scales=cl["scales"].numpy(),
                        rotations=cl["rotations"].numpy(), shs=cl["shs"].numpy())
    b = O.binning(pr, geom)
    assert b["R"] == int(geom["tiles_touched"].sum())
    assert np.all(np.diff(b["keys"].astype(np.uint64)) >= 0)
    tiles = (b["keys"] >> np.uint64(32)).astype(np.int64)
    for t in np.unique(tiles):
        s, e = b["ranges"][t]
        assert np.all(tiles[s:e] == t) and (s == 0 or tiles[s - 1] != t) and (e == b["R"] or tiles[e] != t)
    empty = np.setdiff1d(np.arange(b["ranges"].shape[0]), np.unique(tiles))
    assert np.all(b["ranges"][empty] == 0)
    # ties (same tile, same depth bits) come out in ascending Gaussian index
    same = np.diff(b["keys"].astype(np.uint64)) == 0
    assert np.all(np.diff(b["vals"].astype(np.int64))[same] > 0)
    # depth bits of the key are the Gaussian's depth
    assert np.array_equal((b["keys"] & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                          geom["depths"][b["vals"]].view(np.uint32))


def test_empty_and_culled_inputs():
    cam, cl = small_scene(P=50)
    pr = cam_params(cam, 50, [0.1, 0.2, 0.3])
    behind = cl["means3D"].numpy() * 0 + np.array([[10.0, 4.0, 4.8]], np.float32)  # behind the camera
    fwd = O.render_forward(pr, behind, cl["opacities"].numpy(), scales=cl["scales"].numpy(),
                           rotations=cl["rotations"].numpy(), shs=cl["shs"].numpy())
    assert fwd["bins"]["R"] == 0 and (fwd["geom"]["radii"] == 0).all()
    assert np.allclose(fwd["img"]["color"], np.array([0.1, 0.2, 0.3], np.float32)[:, None, None])
    assert (fwd["img"]["n_contrib"] == 0).all() and (fwd["img"]["final_T"] == 1).all()


# ------------------------------------------------------------------------------ stage-3 binding (golden)
def test_stage3_oracle_matches_reference_golden():
    """oracle/torch_oracle.py stage-3 restatement against tests/golden/binding_stage3.npz (R, scales, cov6 and the
    autograd gradients produced by the reference's own lines)."""
    import torch
    from oracle import torch_oracle as TO
    g = golden("binding_stage3.npz")
    rot = torch.tensor(g["rot_t2w"]).requires_grad_()
    r2 = torch.tensor(g["rotation2d"]).requires_grad_()
    s2 = torch.tensor(g["scaling2d"]).requires_grad_()
    R = TO.stage3_rot_matrix(rot, r2)
    np.testing.assert_allclose(R.detach().numpy(), g["R"], rtol=1e-6, atol=1e-7)
    scales, q = TO.stage3_scales_rotations(rot, s2, r2, float(g["thin_z"]))
    np.testing.assert_allclose(scales.detach().numpy(), g["scales3"], rtol=1e-6)
    # the quaternion reproduces R (pytorch3d's conversion is restated, not executed)
    np.testing.assert_allclose(TO.quat_to_rot(q).detach().numpy(), g["R"], atol=2e-6)
    assert np.allclose(q.detach().norm(dim=1).numpy(), 1.0, atol=1e-6)
    cov = TO.stage3_covariance(rot, s2, r2, float(g["thin_z"]))
    c = g["cov6"].astype(np.float64)
    assert np.linalg.norm(cov.detach().numpy() - c) <= 5e-6 * np.linalg.norm(c)
    (cov * torch.tensor(g["gcov"])).sum().backward()
    for got, name in ((r2.grad, "drotation2d"), (s2.grad, "dscaling2d")):
        ref = g[name]
        assert np.linalg.norm(got.numpy() - ref) <= 1e-4 * np.linalg.norm(ref), name


def test_binning_equals_one_stable_sort_of_the_64_bit_keys():
    """The oracle partitions by tile and merge-sorts every bucket; the result must be THE stable sort of the
    (tile << 32 | depth bits) keys emitted in ascending Gaussian index, row-major over each rectangle
     -- checked against numpy's stable argsort of independently emitted keys."""
    P, W, H = 100_000, 800, 800
    cl = S.random_cloud(P, seed=0)
    cam = S.nerf_synthetic_camera(0, W, H)
    pr = cam_params(cam, P, np.zeros(3, np.float32))
    g = O.preprocess(pr, cl["means3D"].numpy(), cl["opacities"].numpy(), scales=cl["scales"].numpy(),
                     rotations=cl["rotations"].numpy(), shs=cl["shs"].numpy())
    b = O.binning(pr, g)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    f32 = np.float32
    rad, px, py = g["radii"].astype(f32), g["xy"][:, 0], g["xy"][:, 1]
    cl_ = lambda v, hi: np.minimum(np.maximum(v, f32(0)), f32(hi)).astype(np.int64)
    x0, x1 = cl_((px - rad) * f32(0.0625), gx), cl_((px + rad + f32(15)) * f32(0.0625), gx)
    y0, y1 = cl_((py - rad) * f32(0.0625), gy), cl_((py + rad + f32(15)) * f32(0.0625), gy)
    idx = np.nonzero((g["radii"] > 0) & (g["tiles_touched"] > 0))[0]
    w, n = (x1 - x0)[idx], ((x1 - x0) * (y1 - y0))[idx]
    assert np.array_equal(n, g["tiles_touched"][idx].astype(np.int64))
    rep = np.repeat(idx, n)
    off = np.arange(n.sum()) - np.repeat(np.cumsum(n) - n, n)
    ww = np.repeat(w, n)
    ty, tx = np.repeat(y0[idx], n) + off // ww, np.repeat(x0[idx], n) + off % ww
    keys = ((ty * gx + tx).astype(np.uint64) << np.uint64(32)) | np.repeat(g["depths"].view(np.uint32)[idx], n).astype(np.uint64)
    order = np.argsort(keys, kind="stable")
    assert b["R"] == keys.size
    assert np.array_equal(keys[order], b["keys"]) and np.array_equal(rep[order].astype(np.uint32), b["vals"])
    T = gx * gy
    counts = np.bincount((keys >> np.uint64(32)).astype(np.int64), minlength=T)
    ends = np.cumsum(counts)
    ne = counts > 0
    assert np.array_equal(b["ranges"][ne, 0], (ends - counts)[ne]) and np.array_equal(b["ranges"][ne, 1], ends[ne])
    assert not b["ranges"][~ne].any()


# ------------------------------------------------------------------------------ rasteriser oracle: properties
# The splat arithmetic of the oracle has no reference-held vectors, so besides the fp64 autograd
# comparison above it is held to properties any faithful restatement of the published algorithm must have.
def _render_precomp(pr, cl, colors, opac=None, order=None):
    m, s, r = cl["means3D"].numpy(), cl["scales"].numpy(), cl["rotations"].numpy()
    o = cl["opacities"].numpy() if opac is None else opac
    if order is not None:
        m, s, r, o, colors = m[order], s[order], r[order], o[order], colors[order]
    return O.render_forward(pr, np.ascontiguousarray(m), np.ascontiguousarray(o), scales=np.ascontiguousarray(s),
                            rotations=np.ascontiguousarray(r), colors_precomp=np.ascontiguousarray(colors))


def test_oracle_image_is_linear_in_the_colours_and_affine_in_the_background():
    cam, cl = small_scene(P=1500, W=96, H=64, scale=0.06)
    P = 1500
    rng = np.random.default_rng(0)
    c1, c2 = rng.random((P, 3), dtype=np.float32), rng.random((P, 3), dtype=np.float32)
    pr0 = cam_params(cam, P, [0, 0, 0])
    i1 = _render_precomp(pr0, cl, c1)["img"]
    i2 = _render_precomp(pr0, cl, c2)["img"]
    i12 = _render_precomp(pr0, cl, (0.25 * c1 + 1.5 * c2).astype(np.float32))["img"]
    assert np.abs(i12["color"] - (0.25 * i1["color"] + 1.5 * i2["color"])).max() < 2e-5
    # the weights (final T, contributor counts) do not depend on the colours at all
    assert np.array_equal(i1["final_T"], i2["final_T"]) and np.array_equal(i1["n_contrib"], i2["n_contrib"])
    # background: image(bg) = image(0) + final_T * bg, pixel by pixel
    bg = np.array([0.2, 0.7, 0.4], np.float32)
    ib = _render_precomp(cam_params(cam, P, bg.tolist()), cl, c1)["img"]
    assert np.abs(ib["color"] - (i1["color"] + i1["final_T"][None] * bg[:, None, None])).max() < 1e-6
    assert 0.0 < i1["final_T"].min() and i1["final_T"].max() <= 1.0  # T never reaches 0: blending stops at 1e-4


def test_oracle_image_does_not_depend_on_the_input_order_of_the_gaussians():
    cam, cl = small_scene(P=1200, W=96, H=64, scale=0.06, seed=5)
    P = 1200
    rng = np.random.default_rng(1)
    col = rng.random((P, 3), dtype=np.float32)
    pr = cam_params(cam, P, [0.1, 0.1, 0.1])
    a = _render_precomp(pr, cl, col)
    perm = rng.permutation(P)
    b = _render_precomp(pr, cl, col, order=perm)
    # distinct depths (checked): every pixel blends the same Gaussians in the same order, so the bits agree
    assert np.unique(a["geom"]["depths"][a["geom"]["radii"] > 0]).size == int((a["geom"]["radii"] > 0).sum())
    assert np.array_equal(a["img"]["color"], b["img"]["color"])
    assert np.array_equal(a["img"]["final_T"], b["img"]["final_T"]) and np.array_equal(a["img"]["n_contrib"], b["img"]["n_contrib"])
    assert np.array_equal(a["geom"]["radii"][perm], b["geom"]["radii"])


def test_oracle_ignores_gaussians_below_the_alpha_threshold():
    """alpha < 1/255 never contributes (and never counts as a contributor): raising such Gaussians' opacity from
    0 to just under the threshold leaves image and final T bit-identical; pixels are opaque-limited at alpha 0.99."""
    cam, cl = small_scene(P=800, W=96, H=64, scale=0.07, seed=6)
    P = 800
    rng = np.random.default_rng(2)
    col = rng.random((P, 3), dtype=np.float32)
    pr = cam_params(cam, P, [0, 0, 0])
    o = cl["opacities"].numpy().copy()
    ghost = rng.random(P) < 0.4
    o0, o1 = o.copy(), o.copy()
    o0[ghost] = 0.0
    o1[ghost] = 1.0 / 255.0 - 1e-5
    a, b = _render_precomp(pr, cl, col, opac=o0), _render_precomp(pr, cl, col, opac=o1)
    assert np.array_equal(a["img"]["color"], b["img"]["color"]) and np.array_equal(a["img"]["final_T"], b["img"]["final_T"])
    # one fully opaque splat: T after it is 1 - 0.99 at its centre pixel, never 0
    one = {k: v[:1] for k, v in cl.items()}
    c = S.look_at_camera([2.5, 1.0, 1.2], 96, 64, fovx=0.9)
    one["means3D"] = torch.zeros(1, 3)
    one["scales"] = torch.full((1, 3), 0.3)
    f = O.render_forward(cam_params(c, 1, [0, 0, 0]), one["means3D"].numpy(), np.ones((1, 1), np.float32),
                         scales=one["scales"].numpy(), rotations=one["rotations"].numpy(),
                         colors_precomp=np.ones((1, 3), np.float32))
    assert abs(float(f["img"]["final_T"].min()) - 0.01) < 1e-6 and abs(float(f["img"]["color"].max()) - 0.99) < 1e-6


def test_oracle_colour_gradient_is_the_adjoint_of_the_blend():
    """The image is linear in the colours (weights w_i(pixel) = alpha_i T_i), so the blend backward's colour gradient
    must be the transposed map: sum_i <dL/dc_i, c_i> = <dL/dimage, image - final_T * bg>, and the gradient of a second
    colour set through the same geometry is the same linear functional."""
    cam, cl = small_scene(P=1000, W=96, H=64, scale=0.06, seed=7)
    P = 1000
    rng = np.random.default_rng(3)
    col = rng.random((P, 3), dtype=np.float32)
    bg = np.array([0.3, 0.1, 0.6], np.float32)
    pr = cam_params(cam, P, bg.tolist())
    fwd = _render_precomp(pr, cl, col)
    dL = rng.standard_normal((3, 64, 96)).astype(np.float32)
    g = O.blend_backward(pr, fwd["geom"], fwd["bins"], fwd["img"], dL)
    lhs = float((g["dL_dcolor"].astype(np.float64) * col).sum())
    rhs = float((dL.astype(np.float64) * (fwd["img"]["color"] - fwd["img"]["final_T"][None] * bg[:, None, None])).sum())
    assert abs(lhs - rhs) <= 2e-5 * max(abs(rhs), float(np.abs(dL).sum()) * 1e-3)
    # the same functional applied to other colours predicts their image's inner product with dL
    col2 = rng.random((P, 3), dtype=np.float32)
    img2 = _render_precomp(pr, cl, col2)["img"]
    rhs2 = float((dL.astype(np.float64) * (img2["color"] - img2["final_T"][None] * bg[:, None, None])).sum())
    lhs2 = float((g["dL_dcolor"].astype(np.float64) * col2).sum())
    assert abs(lhs2 - rhs2) <= 2e-5 * max(abs(rhs2), float(np.abs(dL).sum()) * 1e-3)
    # Gaussians that reach no pixel get exactly zero
    dead = fwd["geom"]["radii"] == 0
    assert dead.any() and not g["dL_dcolor"][dead].any()
