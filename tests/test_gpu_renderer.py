"""GPU tests of the host-side mirrors: render / render_dyn (dmgs_b200/renderer.py) against the
reference's data flow (python SH -> colors_precomp), and the view-batched accumulation of
dmgs_b200/multiview.py against per-view autograd."""
import math
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from dmgs_b200 import synthetic as S
from oracle import torch_oracle as TO
from util import grad_close

pytestmark = pytest.mark.gpu


def _stage2_scene(F=4000, k=3, seed=2):
    from dmgs_b200.binding import renew_gaussian
    verts, faces = S.jittered_sphere_mesh(F, seed=seed, jitter=0.05)
    bc, rad = S.barycentric_layout(k)
    P = faces.shape[0] * k
    feats = (torch.randn(P, 3, 16, generator=torch.Generator().manual_seed(seed)) * 0.3).cuda().requires_grad_()
    v = verts.cuda().requires_grad_()
    sf = torch.tensor([math.atanh(0.5)], device="cuda", requires_grad=True)
    gs = renew_gaussian(v, faces.cuda(), bc.cuda(), rad, 4.43, sf, feats)
    return gs, v, sf, feats


def test_render_dyn_fused_sh_matches_python_sh():
    """render_dyn with the SH folded into preprocess == the reference's flow
    (gaussian_renderer/__init__.py:166-186: python eval_sh + sigmoid -> colors_precomp)."""
    from dmgs_b200 import GaussianRasterizer
    from dmgs_b200.renderer import render_dyn
    from gpu_util import settings_for
    cam = S.nerf_synthetic_camera(1, 320, 240)
    bg = torch.ones(3, device="cuda")
    pipe = SimpleNamespace(compute_cov3D_python=True, convert_SHs_python=True, debug=False)
    dL = torch.randn(3, 240, 320, generator=torch.Generator().manual_seed(4)).cuda()

    gs, v, sf, feats = _stage2_scene()
    out = render_dyn(cam, gs, pipe, bg)
    assert set(out) == {"render", "viewspace_points", "visibility_filter", "radii"}
    (out["render"] * dL).sum().backward()
    g_fused = [t.grad.clone() for t in (v, sf, feats)]
    assert out["viewspace_points"].grad.shape == (gs["xyz"].shape[0], 3)

    gs2, v2, sf2, feats2 = _stage2_scene()
    campos = cam.camera_center.cuda()
    col = TO.eval_sh_colors(3, feats2.transpose(1, 2), gs2["xyz"], campos, 1)
    ras = GaussianRasterizer(settings_for(cam, (1, 1, 1)))
    img, radii = ras(means3D=gs2["xyz"], means2D=torch.zeros_like(gs2["xyz"]), shs=None, colors_precomp=col,
                     opacities=gs2["opacity"], scales=None, rotations=None, cov3D_precomp=gs2["covariance"])
    (img * dL).sum().backward()
    assert torch.equal(radii, out["radii"])
    assert (img - out["render"]).abs().max().item() <= 1e-5
    for a, b, name in zip(g_fused, (v2.grad, sf2.grad, feats2.grad), ("verts", "scale_factor", "features")):
        grad_close(a.cpu().numpy().reshape(a.shape[0], -1), b.cpu().numpy().reshape(b.shape[0], -1), rtol=2e-4, name=name)


def test_render_stage1_surface():
    from dmgs_b200.renderer import render
    cam = S.nerf_synthetic_camera(0, 200, 160)
    cl = S.random_cloud(3000, seed=1, extent=1.0, log_scale_mean=math.log(0.05))
    t = {k: x.cuda().requires_grad_() for k, x in cl.items()}
    pc = SimpleNamespace(get_xyz=t["means3D"], get_opacity=t["opacities"], get_scaling=t["scales"],
                         get_rotation=t["rotations"], get_features=t["shs"], active_sh_degree=3, max_sh_degree=3)
    pipe = SimpleNamespace(compute_cov3D_python=False, convert_SHs_python=False, debug=False)
    out = render(cam, pc, pipe, torch.zeros(3, device="cuda"))
    out["render"].mean().backward()
    assert out["render"].shape == (3, 160, 200)
    assert out["visibility_filter"].dtype == torch.bool and out["visibility_filter"].sum() > 0
    assert torch.isfinite(t["shs"].grad).all() and t["shs"].grad.abs().sum() > 0
    # sigmoid variant (DMGS's convert_SHs_python flow, __init__.py:74-78) == python sigmoid(eval_sh) -> colors_precomp
    pipe2 = SimpleNamespace(compute_cov3D_python=False, convert_SHs_python=True, debug=False)
    out2 = render(cam, pc, pipe2, torch.zeros(3, device="cuda"))
    col = TO.eval_sh_colors(3, t["shs"].detach(), t["means3D"].detach(), cam.camera_center.cuda(), 1)
    out3 = render(cam, pc, pipe2, torch.zeros(3, device="cuda"), override_color=col)
    assert (out2["render"] - out3["render"]).abs().max().item() <= 1e-5


def test_multiview_accumulate_equals_sum_of_views():
    from dmgs_b200 import GaussianRasterizer, multiview as MV
    from gpu_util import settings_for
    P, W, H, NV = 5000, 160, 120, 3
    cl = S.random_cloud(P, seed=8, extent=1.0, log_scale_mean=math.log(0.05))
    d = {k: v.cuda() for k, v in cl.items()}
    cams = [S.nerf_synthetic_camera(v, W, H) for v in range(NV)]
    sets = [settings_for(c, (0, 0, 0)) for c in cams]
    dLs = [torch.randn(3, H, W, generator=torch.Generator().manual_seed(v)).cuda() for v in range(NV)]
    buf = MV.FlatGradBuffer(P, MV.RASTER_WIDTHS_SH, "cuda")
    vp = MV.ViewParallel()
    inputs = dict(means3D=d["means3D"], opacities=d["opacities"], shs=d["shs"], scales=d["scales"], rotations=d["rotations"])

    def rv(v, acc):
        loss, _, _ = MV.accumulate_view(sets[v], inputs, lambda img: ((img * dLs[v]).sum(), dLs[v]), acc)
        return loss

    loss = vp.step(NV, rv, buf)
    t = {k: x.clone().requires_grad_() for k, x in d.items()}
    tot = 0
    for v in range(NV):
        img, _ = GaussianRasterizer(sets[v])(means3D=t["means3D"], means2D=torch.zeros_like(t["means3D"]), shs=t["shs"],
                                             opacities=t["opacities"], scales=t["scales"], rotations=t["rotations"])
        tot = tot + (img * dLs[v]).sum()
    tot.backward()
    assert abs(loss.item() - tot.item()) <= 1e-4 * abs(tot.item())
    for name in ("means3D", "opacities", "scales", "rotations", "shs"):
        a, b = buf.views[name].cpu().numpy(), t[name].grad.cpu().numpy()
        grad_close(a.reshape(P, -1), b.reshape(P, -1), rtol=2e-4, name=name)


def test_view_streams_equal_single_stream():
    """Views of a step on two CUDA streams (one accumulator each) give the single-stream sum."""
    from dmgs_b200 import multiview as MV
    from gpu_util import settings_for
    P, W, H, NV = 5000, 160, 120, 5
    cl = S.random_cloud(P, seed=9, extent=1.0, log_scale_mean=math.log(0.05))
    d = {k: v.cuda() for k, v in cl.items()}
    sets = [settings_for(S.nerf_synthetic_camera(v, W, H), (0, 0, 0)) for v in range(NV)]
    dLs = [torch.randn(3, H, W, generator=torch.Generator().manual_seed(v)).cuda() for v in range(NV)]
    inputs = dict(means3D=d["means3D"], opacities=d["opacities"], shs=d["shs"], scales=d["scales"], rotations=d["rotations"])
    out = []
    for n in (1, 2):
        vs = MV.ViewStreams(P, MV.RASTER_WIDTHS_SH, torch.device("cuda"), n=n)
        for _ in range(2):  # twice: begin() must reset every accumulator
            vs.begin()
            for v in range(NV):
                vs.run(v, lambda acc, v=v: MV.accumulate_view(sets[v], inputs, lambda img: (None, dLs[v]), acc))
            buf = vs.finish()
        torch.cuda.synchronize()
        out.append({k: x.cpu().numpy().copy() for k, x in buf.views.items()})
    for name in out[0]:
        grad_close(out[1][name].reshape(P, -1), out[0][name].reshape(P, -1), rtol=2e-4, name=name)


@pytest.mark.parametrize("layout", ["PM3", "P3M_sigmoid"])
def test_deferred_sh_gradient_equals_row_accumulation(layout):
    """ViewStreams(deferred_sh_views=...): 16-byte records per view + one dmgs_sh_grad_expand per step give the same
    gradients as the read-modify-write of the SH rows every view (both SH layouts / activations; cameras on a
    close orbit so that part of the cloud is culled in every view)."""
    from dmgs_b200 import multiview as MV
    from gpu_util import settings_for
    P, W, H, NV = 6000, 160, 120, 5
    cl = S.random_cloud(P, seed=10, extent=1.5, log_scale_mean=math.log(0.05))
    d = {k: v.cuda() for k, v in cl.items()}
    cams = [S.look_at_camera([1.6 * math.cos(v), 0.4, 1.6 * math.sin(v)], W, H, fovx=0.8) for v in range(NV)]
    sets = [settings_for(c, (0, 0, 0)) for c in cams]
    dLs = [torch.randn(3, H, W, generator=torch.Generator().manual_seed(v)).cuda() for v in range(NV)]
    if layout == "PM3":
        widths, shs, kw, lay = MV.RASTER_WIDTHS_SH, d["shs"], {}, 0
    else:
        widths = {"means3D": (3,), "means2D": (3,), "opacities": (1,), "scales": (3,), "rotations": (4,), "shs": (3, 16)}
        shs, kw, lay = d["shs"].transpose(1, 2).contiguous(), dict(sh_layout=1, sh_activation=1), 1
    inputs = dict(means3D=d["means3D"], opacities=d["opacities"], shs=shs, scales=d["scales"], rotations=d["rotations"])
    out = []
    for deferred in (0, NV):
        vs = MV.ViewStreams(P, widths, torch.device("cuda"), n=2, deferred_sh_views=deferred)
        for _ in range(2):  # twice: begin() resets the records' bookkeeping and the dense part
            vs.begin()
            for v in range(NV):
                rec = vs.sh_record(v, sets[v].campos)
                vs.run(v, lambda acc, v=v, rec=rec: MV.accumulate_view(sets[v], inputs, lambda img: (None, dLs[v]), acc,
                                                                       sh_record=rec, **kw))
            buf = vs.finish(d["means3D"], shs, 3, sh_layout=lay)
        torch.cuda.synchronize()
        out.append({k: x.cpu().numpy().copy() for k, x in buf.views.items()})
    assert np.abs(out[0]["shs"]).max() > 0
    assert (out[1]["shs"].reshape(P, -1) == 0).all(axis=1).sum() > 0  # some Gaussians are seen by no view
    for name in out[0]:
        grad_close(out[1][name].reshape(P, -1), out[0][name].reshape(P, -1), rtol=2e-4, name=name)


def test_view_batched_training_step_reduces_the_loss():
    """The INTEGRATION.md section 5 sequence end to end: ViewStreams (deferred SH) -> accumulate_view with the fused
    L1+SSIM loss -> finish -> FusedAdam on the flat buffer.  A few steps towards a target rendered from perturbed
    parameters must reduce the loss."""
    from dmgs_b200 import loss_utils as LU, multiview as MV
    from dmgs_b200.optim import FusedAdam
    from dmgs_b200.rasterizer import rasterize_forward
    from gpu_util import settings_for
    P, W, H, NV = 4000, 128, 96, 4
    cl = S.random_cloud(P, seed=12, extent=1.0, log_scale_mean=math.log(0.06))
    sets = [settings_for(S.nerf_synthetic_camera(v, W, H), (0, 0, 0)) for v in range(NV)]
    target = {k: v.cuda() for k, v in cl.items()}
    with torch.no_grad():
        gts = [rasterize_forward(s_, target["means3D"], target["opacities"], target["shs"], None, target["scales"],
                                 target["rotations"], None)[0].clone() for s_ in sets]
    gen = torch.Generator().manual_seed(1)
    params = {k: v.clone() for k, v in target.items()}
    params["shs"] = (params["shs"] + 0.2 * torch.randn(P, 16, 3, generator=gen).cuda()).contiguous()
    params["opacities"] = (params["opacities"] * 0.7).contiguous()
    opt = FusedAdam([{"params": [params[k]], "lr": lr, "name": k} for k, lr in
                     (("means3D", 0.0), ("opacities", 0.02), ("scales", 0.0), ("rotations", 0.0), ("shs", 0.02))],
                    lr=0.0, eps=1e-15)
    vs = MV.ViewStreams(P, MV.RASTER_WIDTHS_SH, torch.device("cuda"), n=2, deferred_sh_views=NV)
    losses = []
    for step in range(6):
        vs.begin()
        per_view = []
        for j in range(NV):
            rec = vs.sh_record(j, sets[j].campos)
            per_view.append(vs.run(j, lambda acc, j=j, rec=rec: MV.accumulate_view(
                sets[j], params, lambda img, j=j: LU.l1_ssim_loss_and_grad(img, gts[j], 0.2), acc, sh_record=rec)[0]))
        vs.finish(params["means3D"], params["shs"], 3)
        vs.all_reduce_(scale=1.0 / NV)
        opt.step(grads=vs.buf.views)
        params["opacities"].clamp_(1e-3, 0.999)
        losses.append(float(torch.stack(per_view).mean()))
    assert losses[-1] < 0.7 * losses[0], losses
    assert all(torch.isfinite(v).all() for v in params.values())


def test_async_binning_matches_sync_and_reports_overflow():
    """configure(async_binning=True): same image / gradients without the host read-back; a frame whose
    instance list outgrows the remembered capacity is never trusted silently: its backward raises
    BinningOverflowError (the module path), check_async() reports it (the training-step path, verify=False), and
    the repeated frame is right."""
    import dmgs_b200
    from dmgs_b200 import GaussianRasterizer
    from dmgs_b200 import rasterizer as RZ
    from gpu_util import settings_for
    P, W, H = 4000, 192, 128
    cl = S.random_cloud(P, seed=11, extent=1.0, log_scale_mean=math.log(0.04))
    d = {k: v.cuda() for k, v in cl.items()}
    cam = S.nerf_synthetic_camera(2, W, H)
    rs = settings_for(cam, (0.3, 0.2, 0.1))

    def run(scales):
        t = {k: v.clone().requires_grad_() for k, v in d.items()}
        ras = GaussianRasterizer(rs)
        img, radii = ras(means3D=t["means3D"], means2D=torch.zeros_like(t["means3D"]), shs=t["shs"],
                         opacities=t["opacities"], scales=scales, rotations=t["rotations"])
        img.square().sum().backward()
        return img.detach(), t["means3D"].grad.clone(), ras.last

    def low_level(scales):
        color, radii, st = RZ.rasterize_forward(rs, d["means3D"], d["opacities"], d["shs"], None, scales, d["rotations"], None)
        g = RZ.rasterize_backward(st, torch.ones_like(color), d["means3D"], d["shs"], scales, d["rotations"], None, False,
                                  verify=False)
        return color, g[0], st

    ref_img, ref_g, _ = run(d["scales"])
    big_img, _, _ = run(d["scales"] * 3.0)
    try:
        dmgs_b200.configure(async_binning=True, capacity_slack=1.25)
        RZ._ASYNC["capacity"].clear()
        img0, g0, st0 = run(d["scales"])  # first frame of this shape: synchronous, learns the capacity
        assert st0.layout_R == st0.num_rendered and dmgs_b200.check_async()
        img1, g1, st1 = run(d["scales"])  # sync-free
        assert st1.layout_R > st1.num_rendered
        assert dmgs_b200.check_async()
        assert torch.equal(img1, ref_img) and torch.equal(img0, ref_img)
        grad_close(g1.cpu().numpy(), ref_g.cpu().numpy(), rtol=2e-4, name="means3D")
        # 3x larger splats: the instance list no longer fits -> the module's backward refuses the frame
        with pytest.raises(RZ.BinningOverflowError):
            run(d["scales"] * 3.0)
        dmgs_b200.check_async()
        img3, _, st3 = run(d["scales"] * 3.0)  # repeated with the raised capacity
        assert dmgs_b200.check_async()
        assert torch.equal(img3, big_img)
        # the training-step path polls once per step instead (verify=False keeps the host running ahead)
        RZ._ASYNC["capacity"].clear()
        low_level(d["scales"])
        assert dmgs_b200.check_async()
        img4, g4, st4 = low_level(d["scales"] * 3.0)
        assert not dmgs_b200.check_async()
        bgimg = torch.tensor([0.3, 0.2, 0.1], device="cuda").view(3, 1, 1).expand_as(img4)
        assert torch.equal(img4, bgimg) and g4.abs().sum() == 0
        img5, _, _ = low_level(d["scales"] * 3.0)
        assert dmgs_b200.check_async() and torch.equal(img5, big_img)
    finally:
        dmgs_b200.configure(async_binning=False)


def test_captured_views_equal_eager_and_report_overflow():
    """ViewStreams.capture / replay: a view's forward + backward as ONE CUDA graph gives the eager step's gradients
    (deferred SH records included), follows in-place parameter updates, and a view whose instance list outgrows the
    capacity baked into its graph is reported by poll_captured() (graph dropped, capture again, repeat the step)."""
    import dmgs_b200
    from dmgs_b200 import multiview as MV
    from dmgs_b200 import rasterizer as RZ
    from gpu_util import settings_for
    P, W, H, NV = 5000, 160, 120, 5
    cl = S.random_cloud(P, seed=12, extent=1.0, log_scale_mean=math.log(0.05))
    d = {k: v.cuda() for k, v in cl.items()}
    sets = [settings_for(S.nerf_synthetic_camera(v, W, H), (0, 0, 0)) for v in range(NV)]
    dLs = [torch.randn(3, H, W, generator=torch.Generator().manual_seed(v)).cuda() for v in range(NV)]
    inputs = dict(means3D=d["means3D"], opacities=d["opacities"], shs=d["shs"], scales=d["scales"], rotations=d["rotations"])
    dev = torch.device("cuda")

    def view(vs, v, acc):
        rec = vs.sh_record(v, sets[v].campos)
        return MV.accumulate_view(sets[v], inputs, lambda img: (None, dLs[v]), acc, sh_record=rec)

    def eager_step(vs):
        vs.begin()
        for v in range(NV):
            vs.run(v, lambda acc, v=v: view(vs, v, acc))
        buf = vs.finish(d["means3D"], d["shs"], 3)
        torch.cuda.synchronize()
        assert dmgs_b200.check_async()
        return {k: x.cpu().numpy().copy() for k, x in buf.views.items()}

    def graph_step(vs):
        vs.begin()
        for v in range(NV):
            if not vs.captured(v):
                vs.capture(v, lambda acc, v=v: view(vs, v, acc))
            vs.replay(v)
        buf = vs.finish(d["means3D"], d["shs"], 3)
        ok = vs.poll_captured()
        torch.cuda.synchronize()
        return ok, {k: x.cpu().numpy().copy() for k, x in buf.views.items()}

    try:
        dmgs_b200.configure(async_binning=True, capacity_slack=1.25)
        RZ._ASYNC["capacity"].clear()
        vs = MV.ViewStreams(P, MV.RASTER_WIDTHS_SH, dev, n=2, deferred_sh_views=NV)
        eager_step(vs)  # learns the capacity, caches the camera values
        ref = eager_step(vs)
        for _ in range(3):  # replays are repeatable; begin() resets the accumulator between them
            ok, got = graph_step(vs)
            assert ok
        assert all(vs.captured(v) for v in range(NV))
        for name in ref:
            grad_close(got[name].reshape(P, -1), ref[name].reshape(P, -1), rtol=2e-4, name=name)
        # parameters updated IN PLACE are what the next replay renders
        d["opacities"].mul_(0.5)
        ref2 = eager_step(vs)
        ok, got2 = graph_step(vs)
        assert ok and np.abs(ref2["shs"] - ref["shs"]).max() > 0
        for name in ref2:
            grad_close(got2[name].reshape(P, -1), ref2[name].reshape(P, -1), rtol=2e-4, name=name)
        # 3x larger splats: the instance lists outgrow the capacity baked into the graphs
        d["scales"].mul_(3.0)
        ok, _ = graph_step(vs)
        assert not ok and not all(vs.captured(v) for v in range(NV))
        ok, got3 = graph_step(vs)  # dropped views are captured again with the raised capacity
        if not ok:  # views that had fitted the first time may need the second raise
            ok, got3 = graph_step(vs)
        assert ok
        ref3 = eager_step(vs)
        for name in ref3:
            grad_close(got3[name].reshape(P, -1), ref3[name].reshape(P, -1), rtol=2e-4, name=name)
    finally:
        dmgs_b200.configure(async_binning=False)


def test_blend_residency_does_not_change_results():
    """dmgs_set_blend_residency: the persistent blend grids hand the 8x8 squares out dynamically; image, final T and
    contributor counts are bit-identical and the gradients agree for every residency (1 CTA/SM: nearly every square
    comes from the device counter; 8: the library default)."""
    from dmgs_b200 import _lib as L
    from dmgs_b200 import rasterizer as RZ
    from gpu_util import settings_for
    P, W, H = 20000, 400, 304
    cl = S.random_cloud(P, seed=13, extent=1.0, log_scale_mean=math.log(0.03))
    d = {k: v.cuda() for k, v in cl.items()}
    rs = settings_for(S.nerf_synthetic_camera(1, W, H), (0.1, 0.2, 0.3))
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(5)).cuda()
    lib = L.lib()
    assert lib.dmgs_set_blend_residency(9, 1) == -7 and lib.dmgs_set_blend_residency(1, -1) == -7
    out = []
    try:
        for k in (8, 3, 1):
            assert lib.dmgs_set_blend_residency(k, k) == 0
            color, radii, st = RZ.rasterize_forward(rs, d["means3D"], d["opacities"], d["shs"], None, d["scales"],
                                                    d["rotations"], None)
            g = RZ.rasterize_backward(st, dL, d["means3D"], d["shs"], d["scales"], d["rotations"], None, False)
            img = st.image[:-256].clone()  # final_T | n_contrib (the last 256 bytes hold the forward's square counter)
            torch.cuda.synchronize()
            out.append((color.clone(), img, [x.cpu().numpy() for x in g if x is not None]))
    finally:
        lib.dmgs_set_blend_residency(8, 8)
    for color, img, g in out[1:]:
        assert torch.equal(color, out[0][0]) and torch.equal(img, out[0][1])  # image; final_T | n_contrib bytes
        for a, b in zip(g, out[0][2]):
            grad_close(a.reshape(P, -1), b.reshape(P, -1), rtol=2e-4, name="grad")


def test_place_smem_budget_does_not_change_the_lists():
    """dmgs_set_place_smem_kb: the placement kernels run with fewer, longer segments; tile ranges and the depth-ordered
    lists are bit-identical for every budget, and the buffer layout does not depend on it (the budget may change
    between a frame's forward and its backward)."""
    from dmgs_b200 import _lib as L
    from dmgs_b200 import rasterizer as RZ
    from gpu_util import settings_for
    P, W, H = 30000, 640, 400
    cl = S.random_cloud(P, seed=14, extent=1.0, log_scale_mean=math.log(0.03))
    d = {k: v.cuda() for k, v in cl.items()}
    rs = settings_for(S.nerf_synthetic_camera(3, W, H), (0.0, 0.0, 0.0))
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(6)).cuda()
    lib = L.lib()
    assert lib.dmgs_set_place_smem_kb(32) == -7 and lib.dmgs_set_place_smem_kb(256) == -7
    out = []
    try:
        for kb_fwd, kb_bwd in ((200, 200), (128, 128), (64, 200), (200, 64)):
            assert lib.dmgs_set_place_smem_kb(kb_fwd) == 0
            color, radii, st = RZ.rasterize_forward(rs, d["means3D"], d["opacities"], d["shs"], None, d["scales"],
                                                    d["rotations"], None)
            b = st.binning_arrays()
            lists = (b["gidx"].clone(), b["ranges"].clone())
            assert lib.dmgs_set_place_smem_kb(kb_bwd) == 0
            g = RZ.rasterize_backward(st, dL, d["means3D"], d["shs"], d["scales"], d["rotations"], None, False)
            torch.cuda.synchronize()
            out.append((color.clone(), lists, [x.cpu().numpy() for x in g if x is not None]))
    finally:
        lib.dmgs_set_place_smem_kb(200)
    assert out[0][1][0].numel() > 50000
    for color, lists, g in out[1:]:
        assert torch.equal(color, out[0][0])
        assert torch.equal(lists[0], out[0][1][0]) and torch.equal(lists[1], out[0][1][1])
        for a, b in zip(g, out[0][2]):
            grad_close(a.reshape(P, -1), b.reshape(P, -1), rtol=2e-4, name="grad")


def test_accumulate_grad_in_place_matches_autograd_accumulation():
    """configure(accumulate_grad_in_place=True): from the second backward of an accumulation loop on, the kernels add into
    the leaves' existing .grad (autograd receives None); the accumulated gradients equal autograd's own accumulation, the
    .grad tensors keep their identity, and inputs without a .grad (a fresh means2D per view) still get theirs."""
    import dmgs_b200
    from dmgs_b200 import GaussianRasterizer
    from gpu_util import settings_for
    P, W, H, NV = 6000, 176, 128, 3
    cl = S.random_cloud(P, seed=15, extent=1.0, log_scale_mean=math.log(0.05))
    sets = [settings_for(S.nerf_synthetic_camera(v, W, H), (0.1, 0.1, 0.1)) for v in range(NV)]
    dLs = [torch.randn(3, H, W, generator=torch.Generator().manual_seed(20 + v)).cuda() for v in range(NV)]
    names = ["means3D", "opacities", "shs", "scales", "rotations"]

    def loop(fused):
        dmgs_b200.configure(accumulate_grad_in_place=fused)
        t = {k: cl[k].cuda().requires_grad_() for k in names}
        ids, m2d_grads = None, []
        for v in range(NV):
            m2d = torch.zeros_like(t["means3D"], requires_grad=True)
            img, _ = GaussianRasterizer(sets[v])(means3D=t["means3D"], means2D=m2d, shs=t["shs"], opacities=t["opacities"],
                                                 scales=t["scales"], rotations=t["rotations"])
            (img * dLs[v]).sum().backward()
            m2d_grads.append(m2d.grad.clone())
            if v == 0:
                ids = {k: t[k].grad.data_ptr() for k in names}
        assert all(t[k].grad.data_ptr() == ids[k] for k in names)
        torch.cuda.synchronize()
        return {k: t[k].grad.cpu().numpy() for k in names}, [g.cpu().numpy() for g in m2d_grads]

    try:
        ref, ref2d = loop(False)
        got, got2d = loop(True)
    finally:
        dmgs_b200.configure(accumulate_grad_in_place=False)
    for k in names:
        grad_close(got[k].reshape(P, -1), ref[k].reshape(P, -1), rtol=2e-4, name=k)
    for a, b in zip(got2d, ref2d):
        assert np.abs(b).max() > 0
        grad_close(a, b, rtol=2e-4, name="means2D")
