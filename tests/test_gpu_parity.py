"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): sort keys, tile ranges, radii, contributor counts bit-exact;
image within 1e-5 absolute; gradients within 1e-4 relative (accumulation order differs)."""
import math

import numpy as np
import pytest
import torch

from dmgs_b200 import synthetic as S
from oracle import oracle as O
from util import cam_params, cov6_from_scale_rot, grad_close, grad_close_conditioned, small_scene

pytestmark = pytest.mark.gpu

MODES = ["sh_scale_rot", "precomp", "sigmoid_features", "sh_deg1_scale_mod", "sh_m4_direct"]


def mode_inputs(mode, cl, P):
    """-> (tensor kwargs, oracle param kwargs, gpu kwargs)"""
    if mode == "sh_scale_rot":
        return dict(scales=cl["scales"], rotations=cl["rotations"], shs=cl["shs"]), {}, {}
    if mode == "sh_deg1_scale_mod":
        return (dict(scales=cl["scales"], rotations=cl["rotations"], shs=cl["shs"]),
                dict(sh_degree=1, scale_modifier=0.7), {})
    if mode == "sh_m4_direct":  # 4 stored coefficients: rows are not staged through TMA (direct loads)
        return (dict(scales=cl["scales"], rotations=cl["rotations"], shs=cl["shs"][:, :4, :].contiguous()),
                dict(sh_degree=1, sh_coeffs=4), {})
    if mode == "precomp":
        col = torch.rand(P, 3, generator=torch.Generator().manual_seed(9))
        return dict(cov3D_precomp=cov6_from_scale_rot(cl["scales"], cl["rotations"]), colors_precomp=col), {}, {}
    feats = cl["shs"].transpose(1, 2).contiguous()
    return (dict(scales=cl["scales"], rotations=cl["rotations"], shs=feats), dict(sh_layout=1, sh_act=1, sh_degree=3),
            dict(sh_layout=1, sh_activation=1))


def run_both(cam, cl, mode, bg=(0.2, 0.5, 0.7)):
    from gpu_util import gpu_forward, settings_for, state_numpy
    P = cl["means3D"].shape[0]
    inp, pk, gk = mode_inputs(mode, cl, P)
    pr = cam_params(cam, P, np.array(bg, np.float32), **pk)
    npin = {k: v.numpy() for k, v in inp.items()}
    ref = O.render_forward(pr, cl["means3D"].numpy(), cl["opacities"].numpy(), **npin)
    st_ = settings_for(cam, bg, sh_degree=pk.get("sh_degree", 3), scale_modifier=pk.get("scale_modifier", 1.0))
    color, radii, st = gpu_forward(st_, cl["means3D"], cl["opacities"], **gk, **inp)
    return pr, npin, ref, color, radii, st, state_numpy(st)


def assert_forward_parity(ref, color, radii, st, arrays):
    g, b, im, keys = arrays
    rg, rb, ri = ref["geom"], ref["bins"], ref["img"]
    vis = rg["radii"] > 0
    assert np.array_equal(radii.cpu().numpy(), rg["radii"]), "radii"
    assert np.array_equal(g["tiles_touched"].view(np.uint32), rg["tiles_touched"]), "tiles_touched"
    assert np.array_equal(g["depths"].view(np.uint32)[vis], rg["depths"].view(np.uint32)[vis]), "depth bits"
    assert np.array_equal(g["rec"][vis, 0:2].view(np.uint32), rg["xy"][vis].view(np.uint32)), "xy bits"
    assert np.array_equal(g["rec"][vis, 2:6].view(np.uint32), rg["conic_opacity"][vis].view(np.uint32)), "conic bits"
    assert np.array_equal(g["rgb"][vis, :3].view(np.uint32), rg["rgb"][vis].view(np.uint32)), "rgb bits"
    assert st.num_rendered == rb["R"], "num_rendered"
    assert np.array_equal(keys, rb["keys"]), "sorted 64-bit keys"
    assert np.array_equal(b["gidx"].view(np.uint32), rb["vals"]), "sorted values"
    assert np.array_equal(b["ranges"].view(np.uint32), rb["ranges"]), "tile ranges"
    assert np.array_equal(im["n_contrib"].view(np.uint32), ri["n_contrib"]), "n_contrib"
    assert np.array_equal(im["final_T"].view(np.uint32), ri["final_T"].view(np.uint32)), "final_T bits"
    assert np.abs(color.cpu().numpy() - ri["color"]).max() <= 1e-5, "image"


def test_exp_bit_exact():
    from dmgs_b200 import _lib as L
    x = torch.cat([torch.linspace(-90, 5, 500001), torch.tensor([-1e30, 0.0, -0.0, 100.0, float("nan")])]).cuda()
    y = torch.empty_like(x)
    L.check(L.lib().dmgs_exp_array(L.ptr(x), L.ptr(y), x.numel(), None), "exp")
    torch.cuda.synchronize()
    ref = O.exp_array(x.cpu().numpy())
    assert np.array_equal(y.cpu().numpy().view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("mode", MODES)
def test_forward_parity_small(mode):
    cam, cl = small_scene(P=3000, W=200, H=136, scale=0.05)
    _, _, ref, color, radii, st, arrays = run_both(cam, cl, mode)
    assert ref["bins"]["R"] > 3000
    assert_forward_parity(ref, color, radii, st, arrays)


def test_forward_parity_c1_100k():
    cam = S.nerf_synthetic_camera(0)
    cl = S.random_cloud(100_000, seed=0)
    _, _, ref, color, radii, st, arrays = run_both(cam, cl, "sh_scale_rot", bg=(0, 0, 0))
    assert_forward_parity(ref, color, radii, st, arrays)


def test_forward_parity_radix_tile_partition(monkeypatch):
    """Images above PLACE_MAX_TILES tiles use the radix tile partition instead of direct placement;
    DMGS_TILE_PARTITION=radix forces that path at any size."""
    monkeypatch.setenv("DMGS_TILE_PARTITION", "radix")
    cam, cl = small_scene(P=3000, W=200, H=136, scale=0.05)
    for mode in ("sh_scale_rot", "precomp"):
        _, _, ref, color, radii, st, arrays = run_both(cam, cl, mode)
        assert_forward_parity(ref, color, radii, st, arrays)
    cam, cl = small_scene(P=400, W=333, H=250, scale=0.5)
    _, _, ref, color, radii, st, arrays = run_both(cam, cl, "sh_scale_rot")
    assert_forward_parity(ref, color, radii, st, arrays)


def test_forward_parity_big_splats_and_ties():
    # large Gaussians (rectangles of hundreds of tiles -> warp-cooperative emission) and exact
    # duplicates (same tile, bit-identical depth -> order must be ascending Gaussian index)
    cam, cl = small_scene(P=400, W=333, H=250, scale=0.5)
    for k in cl:
        cl[k] = torch.cat([cl[k], cl[k][:150]], 0).contiguous()
    _, _, ref, color, radii, st, arrays = run_both(cam, cl, "sh_scale_rot")
    assert ref["geom"]["tiles_touched"].max() > 64
    assert_forward_parity(ref, color, radii, st, arrays)


@pytest.mark.parametrize("mode", MODES)
def test_backward_parity(mode):
    from gpu_util import settings_for
    from dmgs_b200.rasterizer import rasterize_backward
    cam, cl = small_scene(P=3000, W=200, H=136, scale=0.05)
    pr, npin, ref, color, radii, st, arrays = run_both(cam, cl, mode)
    H, W = cam.image_height, cam.image_width
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(5))
    bw = O.render_backward(pr, ref, dL.numpy(), cl["means3D"].numpy(), scales=npin.get("scales"),
                           rotations=npin.get("rotations"), shs=npin.get("shs"), precomp_color="colors_precomp" in npin,
                           abs_sums=True)
    d = lambda k: None if k not in npin else torch.tensor(npin[k]).cuda()
    g = rasterize_backward(st, dL.cuda(), cl["means3D"].cuda(), d("shs"), d("scales"), d("rotations"),
                           d("cov3D_precomp"), "colors_precomp" in npin)
    torch.cuda.synchronize()
    g_means3D, g_means2D, g_shs, g_col, g_op, g_scales, g_rots, g_cov = [None if t is None else t.cpu().numpy() for t in g]
    grad_close(g_means3D, bw["dL_dmeans3D"], name="means3D")
    grad_close(g_means2D[:, :2], bw["dL_dmean2D"], name="means2D")
    assert (g_means2D[:, 2] == 0).all()
    # the two pure blend-backward outputs, element by element against their conditioning (tests/util.py)
    grad_close_conditioned(g_op[:, 0], bw["dL_dopacity"], bw["abs9"][:, 5], name="opacity")
    grad_close_conditioned(g_means2D[:, :2], bw["dL_dmean2D"], bw["abs9"][:, 0:2], name="means2D (conditioned)")
    if g_shs is not None:
        grad_close(g_shs, bw["dL_dshs"], name="shs")
    if g_col is not None:
        grad_close(g_col, bw["dL_dcolors_precomp"], name="colors_precomp")
    if g_scales is not None:
        grad_close(g_scales, bw["dL_dscales"], name="scales")
        grad_close(g_rots, bw["dL_drotations"], name="rotations")
    if g_cov is not None:
        grad_close(g_cov, bw["dL_dcov3D"], name="cov3D")
    for t in (g_means3D, g_means2D, g_op):
        assert np.isfinite(t).all()


def test_backward_accumulate_mode():
    """accumulate=1 adds into existing gradient buffers (TMA reduce-add for the SH rows)."""
    from dmgs_b200.rasterizer import rasterize_backward
    cam, cl = small_scene(P=3000, W=200, H=136, scale=0.05)
    for mode in ("sh_scale_rot", "sigmoid_features", "sh_m4_direct", "precomp"):
        pr, npin, ref, color, radii, st, arrays = run_both(cam, cl, mode)
        dL = torch.randn(3, cam.image_height, cam.image_width, generator=torch.Generator().manual_seed(5)).cuda()
        d = lambda k: None if k not in npin else torch.tensor(npin[k]).cuda()
        args = (st, dL, cl["means3D"].cuda(), d("shs"), d("scales"), d("rotations"), d("cov3D_precomp"), "colors_precomp" in npin)
        g1 = rasterize_backward(*args)
        P = 3000
        acc = {"means3D": torch.zeros(P, 3).cuda(), "means2D": torch.zeros(P, 3).cuda(), "opacities": torch.zeros(P, 1).cuda(),
               "colors_precomp": torch.zeros(P, 3).cuda(), "scales": torch.zeros(P, 3).cuda(), "rotations": torch.zeros(P, 4).cuda(),
               "cov3D_precomp": torch.zeros(P, 6).cuda()}
        if "shs" in npin:
            acc["shs"] = torch.zeros_like(d("shs"))
        acc["means2D"][:, 2] = 7.0  # never touched
        rasterize_backward(*args, accumulate_into=acc)
        rasterize_backward(*args, accumulate_into=acc)
        torch.cuda.synchronize()
        names = ["means3D", "means2D", "shs", "colors_precomp", "opacities", "scales", "rotations", "cov3D_precomp"]
        for n, g in zip(names, g1):
            if g is None:
                continue
            got = acc[n] * 0.5  # two accumulations
            if n == "means2D":
                assert torch.all(acc[n][:, 2] == 7.0)
                got[:, 2] = 0.0
            # the blend backward's atomics make every run's summation order different
            grad_close(got.cpu().numpy(), g.cpu().numpy(), rtol=2e-4, name=f"{mode}/{n}")


def test_module_autograd_contract():
    """The nn.Module surface: kwargs call, (image, radii) return, means2D.grad [P,3], errors."""
    from gpu_util import settings_for
    from dmgs_b200 import GaussianRasterizer
    cam, cl = small_scene(P=2000, W=160, H=96, scale=0.05)
    P = 2000
    rs = settings_for(cam, (1, 1, 1))
    ras = GaussianRasterizer(raster_settings=rs)
    means3D = cl["means3D"].cuda().requires_grad_()
    means2D = torch.zeros_like(means3D, requires_grad=True)
    means2D.retain_grad()
    op = cl["opacities"].cuda().requires_grad_()
    sc, rot, shs = cl["scales"].cuda().requires_grad_(), cl["rotations"].cuda().requires_grad_(), cl["shs"].cuda().requires_grad_()
    img, radii = ras(means3D=means3D, means2D=means2D, shs=shs, colors_precomp=None, opacities=op, scales=sc,
                     rotations=rot, cov3D_precomp=None)
    assert img.shape == (3, 96, 160) and radii.shape == (P,) and radii.dtype == torch.int32
    img.square().mean().backward()
    assert means2D.grad.shape == (P, 3) and torch.all(means2D.grad[:, 2] == 0)
    for t in (means3D, op, sc, rot, shs):
        assert t.grad is not None and torch.isfinite(t.grad).all()
    assert (means2D.grad[radii == 0] == 0).all()
    vis = ras.markVisible(means3D.detach())
    assert vis.dtype == torch.bool and torch.all(vis[radii > 0])
    with pytest.raises(Exception):
        ras(means3D=means3D, means2D=means2D, shs=shs, colors_precomp=torch.zeros(P, 3).cuda(), opacities=op,
            scales=sc, rotations=rot, cov3D_precomp=None)
    with pytest.raises(Exception):
        ras(means3D=means3D, means2D=means2D, shs=shs, colors_precomp=None, opacities=op, scales=None,
            rotations=None, cov3D_precomp=None)


def test_empty_and_culled():
    from gpu_util import settings_for
    from dmgs_b200 import GaussianRasterizer
    cam, cl = small_scene(P=64)
    rs = settings_for(cam, (0.1, 0.2, 0.3))
    ras = GaussianRasterizer(raster_settings=rs)
    behind = (cl["means3D"] * 0 + torch.tensor([[10.0, 4.0, 4.8]])).cuda().requires_grad_()
    img, radii = ras(means3D=behind, means2D=torch.zeros_like(behind), shs=cl["shs"].cuda(), opacities=cl["opacities"].cuda(),
                     scales=cl["scales"].cuda(), rotations=cl["rotations"].cuda())
    assert (radii == 0).all()
    assert torch.allclose(img, torch.tensor([0.1, 0.2, 0.3]).cuda().view(3, 1, 1).expand_as(img))
    img.sum().backward()
    assert torch.all(behind.grad == 0)
    e = torch.zeros(0, 3).cuda()
    img0, radii0 = ras(means3D=e, means2D=e, shs=torch.zeros(0, 16, 3).cuda(), opacities=torch.zeros(0, 1).cuda(),
                       scales=e, rotations=torch.zeros(0, 4).cuda())
    assert radii0.numel() == 0 and torch.allclose(img0, img)


@pytest.mark.parametrize("k", [1, 3, 6])
@pytest.mark.parametrize("adaptive", [True, False])
def test_binding_parity(k, adaptive):
    from dmgs_b200.binding import bind_faces
    verts, faces = S.jittered_sphere_mesh(5000, seed=k, jitter=0.1)
    bc, rad = S.barycentric_layout(k)
    v = verts.cuda().requires_grad_()
    sf = torch.tensor([0.37], device="cuda", requires_grad=True)
    g = float((torch.tanh(sf) * 2.0).item())  # the same fp32 value the kernel reads
    ref = O.bind_forward(verts.numpy(), faces.numpy(), bc.numpy(), rad, 4.43e-6, g, adaptive)
    xyz, cov6 = bind_faces(v, faces.cuda(), bc.cuda(), rad, 4.43e-6, sf, max_scale=2.0, adaptive_cov=adaptive)
    assert np.array_equal(xyz.detach().cpu().numpy().view(np.uint32), ref["xyz"].view(np.uint32))
    assert np.array_equal(cov6.detach().cpu().numpy().view(np.uint32), ref["cov6"].view(np.uint32))
    gen = torch.Generator().manual_seed(3)
    gx, gc = torch.randn(xyz.shape, generator=gen), torch.randn(cov6.shape, generator=gen) * 1e3
    ((xyz * gx.cuda()).sum() + (cov6 * gc.cuda()).sum()).backward()
    bw = O.bind_backward(verts.numpy(), faces.numpy(), bc.numpy(), rad, 4.43e-6, g, gx.numpy(), gc.numpy(), adaptive)
    grad_close(v.grad.cpu().numpy(), bw["dverts"], rtol=2e-4, name="dverts")
    dsf = bw["dg"] * 2 * (1 - math.tanh(0.37) ** 2)
    assert abs(sf.grad.item() - dsf) <= 2e-4 * abs(dsf)


def test_binding_golden_on_gpu():
    from dmgs_b200.binding import bind_faces
    from util import golden
    g = golden("binding_stage2.npz")
    for k in (1, 3, 6):
        tag = f"k{k}_adp"
        bc, rad = S.barycentric_layout(k)
        v = torch.tensor(g[f"{tag}_verts"]).cuda().requires_grad_()
        sf = torch.tensor(g[f"{tag}_scale_factor"]).cuda().requires_grad_()
        xyz, cov6 = bind_faces(v, torch.tensor(g[f"{tag}_faces"]).cuda(), bc.cuda(), rad, 4.43 * 1e-6, sf)
        assert np.allclose(xyz.detach().cpu().numpy(), g[f"{tag}_xyz"], rtol=1e-5, atol=1e-6)
        c = cov6.detach().cpu().numpy().astype(np.float64)
        assert np.linalg.norm(c - g[f"{tag}_cov6"]) <= 5e-6 * np.linalg.norm(g[f"{tag}_cov6"])
        ((xyz * torch.tensor(g[f"{tag}_gxyz"]).cuda()).sum() + (cov6 * torch.tensor(g[f"{tag}_gcov"]).cuda()).sum()).backward()
        ref = g[f"{tag}_dverts"]
        assert np.linalg.norm(v.grad.cpu().numpy() - ref) <= 1e-4 * np.linalg.norm(ref)
        # face_normals(unit=True) (geo/mesh_utils.py:43-57) = third column of the reference's rot_t2w
        from dmgs_b200.binding import face_normals
        n = face_normals(v.detach(), torch.tensor(g[f"{tag}_faces"]).cuda(), unit=True)
        assert np.allclose(n.cpu().numpy(), g[f"{tag}_rot_t2w"][:, :, 2], rtol=1e-5, atol=1e-6)
        assert abs(sf.grad.item() - g[f"{tag}_dscale_factor"][0]) <= 1e-4 * abs(g[f"{tag}_dscale_factor"][0])


# ------------------------------------------------------------------------------ stage-3 binding
def test_stage3_golden_and_oracle_on_gpu():
    """dmgs_stage3_forward/backward (+ bind_frame) against tests/golden/binding_stage3.npz (the reference's own lines:
    R, scales, cov6 and the gradients to verts / _rotation / _scaling) and, for the quaternion path, against the
    oracle's torch restatement with autograd."""
    from dmgs_b200.binding import bind_frame, stage3_covariance, stage3_scales_rotations
    from oracle import torch_oracle as TO
    from util import golden
    g = golden("binding_stage3.npz")
    bc, _ = S.barycentric_layout(3)
    v = torch.tensor(g["verts"]).cuda().requires_grad_()
    r2 = torch.tensor(g["rotation2d"]).cuda().requires_grad_()
    s2 = torch.tensor(g["scaling2d"]).cuda().requires_grad_()
    xyz, rot = bind_frame(v, torch.tensor(g["faces"]).cuda(), bc.cuda())
    cov = stage3_covariance(rot, s2, r2, float(g["thin_z"]))
    assert np.allclose(xyz.detach().cpu().numpy(), g["xyz"], rtol=1e-5, atol=1e-6)
    c = g["cov6"].astype(np.float64)
    assert np.linalg.norm(cov.detach().cpu().numpy() - c) <= 5e-6 * np.linalg.norm(c)
    ((xyz * torch.tensor(g["gxyz"]).cuda()).sum() + (cov * torch.tensor(g["gcov"]).cuda()).sum()).backward()
    for got, name in ((v.grad, "dverts"), (r2.grad, "drotation2d"), (s2.grad, "dscaling2d")):
        ref = g[name]
        assert np.linalg.norm(got.cpu().numpy() - ref) <= 1e-4 * np.linalg.norm(ref), name
    # (scales, quaternion) path: forward vs golden R / scales, gradients vs the oracle's autograd
    gen = torch.Generator().manual_seed(21)
    P = r2.shape[0]
    gs, gq = torch.randn(P, 3, generator=gen), torch.randn(P, 4, generator=gen)
    rot_c = torch.tensor(g["rot_t2w"])
    o_rot, o_r2, o_s2 = (t.clone().double().requires_grad_() for t in (rot_c, torch.tensor(g["rotation2d"]), torch.tensor(g["scaling2d"])))
    os_, oq = TO.stage3_scales_rotations(o_rot, o_s2, o_r2, float(g["thin_z"]))
    ((os_ * gs.double()).sum() + (oq * gq.double()).sum()).backward()
    d_rot, d_r2, d_s2 = (t.clone().cuda().requires_grad_() for t in (rot_c, torch.tensor(g["rotation2d"]), torch.tensor(g["scaling2d"])))
    sc, q = stage3_scales_rotations(d_rot, d_s2, d_r2, float(g["thin_z"]))
    np.testing.assert_allclose(sc.detach().cpu().numpy(), g["scales3"], rtol=2e-6)
    np.testing.assert_allclose(q.detach().cpu().numpy(), oq.detach().float().numpy(), atol=2e-6)
    np.testing.assert_allclose(TO.quat_to_rot(q.detach().cpu()).numpy(), g["R"], atol=2e-6)
    ((sc * gs.cuda()).sum() + (q * gq.cuda()).sum()).backward()
    for got, ref, name in ((d_rot.grad, o_rot.grad, "rot_t2w"), (d_r2.grad, o_r2.grad, "rotation2d"), (d_s2.grad, o_s2.grad, "scaling2d")):
        ref = ref.numpy()
        assert np.linalg.norm(got.cpu().numpy() - ref) <= 1e-4 * np.linalg.norm(ref), name


def test_stage3_all_quaternion_branches_and_large():
    """Frames covering all four matrix_to_quaternion candidates (rotations by > 120 degrees about each axis) and a
    300 k-Gaussian run against the oracle."""
    from dmgs_b200.binding import stage3_covariance, stage3_scales_rotations
    from oracle import torch_oracle as TO
    gen = torch.Generator().manual_seed(8)
    F, k = 50_000, 6
    q0 = torch.nn.functional.normalize(torch.randn(F, 4, generator=gen), dim=1)
    rot = TO.quat_to_rot(q0).contiguous()                      # random proper rotations as face frames
    r2 = torch.randn(F * k, 2, generator=gen)
    s2 = torch.randn(F * k, 2, generator=gen) * 0.3 - 3.0
    ref_s, ref_q = TO.stage3_scales_rotations(rot, s2, r2, 4.43e-6)
    R = TO.stage3_rot_matrix(rot, r2)
    t2 = torch.stack([1 + R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2], 1 + R[:, 0, 0] - R[:, 1, 1] - R[:, 2, 2],
                      1 - R[:, 0, 0] + R[:, 1, 1] - R[:, 2, 2], 1 - R[:, 0, 0] - R[:, 1, 1] + R[:, 2, 2]], 1)
    assert set(t2.argmax(1).tolist()) == {0, 1, 2, 3}
    sc, q = stage3_scales_rotations(rot.cuda(), s2.cuda(), r2.cuda(), 4.43e-6)
    np.testing.assert_allclose(sc.cpu().numpy(), ref_s.numpy(), rtol=2e-6)
    # a near tie between two candidates may pick the other one: same rotation, possibly the opposite sign
    qc = q.cpu()
    sign = torch.sign((qc * ref_q).sum(1, keepdim=True))
    assert float((qc * sign - ref_q).abs().max()) <= 5e-6
    cov = stage3_covariance(rot.cuda(), s2.cuda(), r2.cuda(), 4.43e-6).cpu().double()
    ref_c = TO.stage3_covariance(rot.double(), s2.double(), r2.double(), 4.43e-6)
    assert float(torch.linalg.norm(cov - ref_c) / torch.linalg.norm(ref_c)) <= 5e-6
