"""GPU parity tests of the rows next to the hot path, through the C ABI:
fused L1+SSIM loss, frustum test + face compaction, fused Adam -- against the golden vectors produced
by the reference's own Python, against the CPU oracle on seeded inputs, and at full size through
size-independent properties."""
import os

import numpy as np
import pytest
import torch

from oracle import next_rows as N

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------------------ L1 + SSIM
@pytest.mark.parametrize("tag", ["a", "b"])
def test_loss_matches_reference_golden(tag):
    from dmgs_b200 import loss_utils as LU
    g = np.load(os.path.join(GOLD, "loss_l1_ssim.npz"))
    img = torch.tensor(g[f"img_{tag}"]).cuda().requires_grad_()
    gt = torch.tensor(g[f"gt_{tag}"]).cuda()
    l1, ss = LU.l1_loss(img, gt), LU.ssim(img, gt)
    assert abs(l1.item() - g[f"l1_{tag}"]) <= 1e-6       # tolerance: 1e-5 absolute on image-level quantities
    assert abs(ss.item() - g[f"ssim_{tag}"]) <= 1e-5
    loss = LU.l1_ssim_loss(img, gt, 0.2)
    assert abs(loss.item() - g[f"loss_{tag}"]) <= 1e-5
    loss.backward()
    ref = g[f"grad_{tag}"]
    got = img.grad.cpu().numpy()
    assert np.linalg.norm(got - ref) <= 1e-4 * np.linalg.norm(ref)  # gradients: 1e-4 relative
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()
    l2, g2 = LU.l1_ssim_loss_and_grad(img.detach(), gt, 0.2)  # the no-autograd path gives the same numbers
    assert abs(l2.item() - loss.item()) <= 1e-7 and torch.allclose(g2, img.grad, rtol=1e-6, atol=1e-12)
    # separate pieces: the L1 gradient is exact (sign / N, zero on exact ties)
    img.grad = None
    LU.l1_loss(img, gt).backward()
    np.testing.assert_allclose(img.grad.cpu().numpy(), g[f"grad_l1_{tag}"], rtol=1e-6, atol=0)
    img.grad = None
    LU.ssim(img, gt).backward()
    ref = g[f"grad_ssim_{tag}"]
    assert np.abs(img.grad.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()


def test_ssim_batched_matches_reference_golden():
    from dmgs_b200 import loss_utils as LU
    g = np.load(os.path.join(GOLD, "loss_l1_ssim.npz"))
    got = LU.ssim(torch.tensor(g["img_batch"]).cuda(), torch.tensor(g["gt_batch"]).cuda(), size_average=False)
    np.testing.assert_allclose(got.cpu().numpy(), g["ssim_batch"], atol=1e-5)


def test_loss_matches_oracle_ragged_and_full_size():
    from dmgs_b200 import loss_utils as LU
    gen = torch.Generator().manual_seed(2)
    for (H, W) in [(1, 1), (5, 70), (33, 32), (97, 131)]:  # smaller than the window, not multiples of the tile
        gt = torch.rand(3, H, W, generator=gen)
        img = (gt + 0.3 * torch.randn(3, H, W, generator=gen)).clamp(0, 1)
        l1, ss, dl1, dss = N.l1_ssim(img.numpy(), gt.numpy())
        x = img.cuda().requires_grad_()
        loss = LU.l1_ssim_loss(x, gt.cuda(), 0.2)
        assert abs(loss.item() - (0.8 * l1 + 0.2 * (1 - ss))) <= 1e-5, (H, W)
        loss.backward()
        ref = 0.8 * dl1 - 0.2 * dss
        assert np.abs(x.grad.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max(), (H, W)
    # full size (800x800): identical images -> ssim == 1, l1 == 0, zero L1 gradient; and the loss is
    # symmetric in its arguments
    a = torch.rand(3, 800, 800, generator=gen).cuda()
    b = torch.rand(3, 800, 800, generator=gen).cuda()
    assert abs(LU.ssim(a, a).item() - 1.0) <= 1e-6 and LU.l1_loss(a, a).item() == 0.0
    assert abs(LU.l1_ssim_loss(a, b).item() - LU.l1_ssim_loss(b, a).item()) <= 1e-6
    x = a.clone().requires_grad_()
    LU.l1_ssim_loss(x, b, 0.2).backward()
    # directional derivative check in fp64-free form: loss(x + t d) - loss(x - t d) ~ 2 t <grad, d>
    d = torch.randn(3, 800, 800, generator=torch.Generator().manual_seed(4)).cuda()
    t = 1e-2
    num = (LU.l1_ssim_loss(a + t * d, b, 1.0).item() - LU.l1_ssim_loss(a - t * d, b, 1.0).item()) / (2 * t)
    y = a.clone().requires_grad_()
    LU.l1_ssim_loss(y, b, 1.0).backward()
    ana = (y.grad * d).sum().item()
    assert abs(num - ana) <= 2e-2 * abs(ana) + 1e-7


# ------------------------------------------------------------------------------ frustum
def test_frustum_matches_reference_golden():
    from dmgs_b200 import frustum as FR
    g = np.load(os.path.join(GOLD, "frustum.npz"))
    proj = torch.tensor(g["proj"]).cuda()
    verts, faces = torch.tensor(g["verts"]).cuda(), torch.tensor(g["faces"]).cuda()
    mask, gs_mask, vis = FR.cull_faces(proj, verts, faces, gs_per_face=3)
    assert np.array_equal(mask.cpu().numpy(), g["face_mask"])
    assert np.array_equal(vis.cpu().numpy(), g["faces_visible"])
    assert np.array_equal(gs_mask.cpu().numpy(), np.repeat(g["face_mask"], 3))
    assert np.array_equal(FR.in_frustum(proj, verts).cpu().numpy(), g["vert_mask"])
    grid = torch.tensor(g["grid"]).cuda()
    for pid, npc in [(-1, 1), (0, 2), (1, 2), (0, 4), (1, 4), (2, 4), (3, 4)]:
        m = FR.in_frustum(proj, grid, float(g["cube_len"]), pid, npc)
        assert np.array_equal(m.cpu().numpy(), g[f"grid_mask_{pid}_{npc}"]), (pid, npc)
    with pytest.raises(NotImplementedError):
        FR.in_frustum(proj, grid, 0.1, 2, 2)


def test_frustum_matches_oracle_large_and_edge_cases():
    from dmgs_b200 import frustum as FR, synthetic as S
    cam = S.look_at_camera([0.8, 0.3, 1.1], 640, 400, fovx=0.7)
    verts, faces = S.jittered_sphere_mesh(300_000, seed=6, jitter=0.05)
    proj = cam.full_proj_transform
    ref = N.in_frustum(proj.numpy(), verts.numpy(), faces=faces.numpy())
    mask, _, vis = FR.cull_faces(proj.cuda(), verts.cuda(), faces.cuda())
    assert 0 < ref.sum() < ref.size
    assert np.array_equal(mask.cpu().numpy(), ref)
    assert np.array_equal(vis.cpu().numpy(), faces.numpy()[ref])      # order preserved, like faces[mask]
    # nothing visible / everything visible / no faces
    far = verts + torch.tensor([100.0, 0.0, 0.0])
    m0, _, v0 = FR.cull_faces(proj.cuda(), far.cuda(), faces.cuda())
    assert not m0.any() and v0.shape == (0, 3)
    tiny = verts * 1e-3
    m1, _, v1 = FR.cull_faces(proj.cuda(), tiny.cuda(), faces.cuda())
    assert m1.all() and torch.equal(v1, faces.cuda())
    m2, _, v2 = FR.cull_faces(proj.cuda(), verts.cuda(), faces[:0].cuda())
    assert m2.numel() == 0 and v2.shape == (0, 3)


# ------------------------------------------------------------------------------ Adam
def test_adam_matches_torch_golden():
    from dmgs_b200.optim import FusedAdam
    g = np.load(os.path.join(GOLD, "adam.npz"))
    names, lrs = [str(n) for n in g["names"]], [float(v) for v in g["lrs"]]
    params = {k: torch.tensor(g[f"p0_{k}"]).cuda() for k in names}
    opt = FusedAdam([{"params": [params[k]], "lr": lr, "name": k} for k, lr in zip(names, lrs)], lr=0.0, eps=1e-15)
    for t in range(int(g["steps"])):
        grads = {k: torch.tensor(g[f"g{t}_{k}"]).cuda() for k in names}
        opt.step(grads=grads, zero_grad=True)
        for k, lr in zip(names, lrs):
            ref = g[f"p{t + 1}_{k}"]
            got = params[k].cpu().numpy()
            assert np.all(np.abs(got - ref) <= 2e-6 * (np.abs(ref) + lr)), (k, t)
            assert not grads[k].any()  # zero_grad folded into the step
    for k in names:
        st = opt.state[params[k]]
        assert np.abs(st["exp_avg"].cpu().numpy() - g[f"m_{k}"]).max() <= 2e-6 * np.abs(g[f"m_{k}"]).max()
        assert np.abs(st["exp_avg_sq"].cpu().numpy() - g[f"v_{k}"]).max() <= 2e-6 * np.abs(g[f"v_{k}"]).max()


def test_adam_flat_buffer_split_lr_and_scale_vs_torch():
    """One launch over a FlatGradBuffer: the [P,16,3] SH tensor with different DC / rest learning rates
    (scene/gaussian_model.py training_setup: feature_lr and feature_lr / 20), view averaging folded in."""
    from dmgs_b200 import multiview as MV
    from dmgs_b200.optim import FusedAdam
    P, NV = 100_003, 8
    gen = torch.Generator().manual_seed(9)
    buf = MV.FlatGradBuffer(P, MV.RASTER_WIDTHS_SH, "cuda")
    lrs = {"means3D": 1.6e-4, "opacities": 0.05, "scales": 0.005, "rotations": 0.001}
    params = {k: torch.randn(P, *w, generator=gen).cuda() for k, w in MV.RASTER_WIDTHS_SH.items() if k != "means2D"}
    groups = [{"params": [params[k]], "lr": lrs[k], "name": k} for k in lrs]
    groups.append({"params": [params["shs"]], "lr": 0.0025, "lr_hi": 0.0025 / 20, "period": 48, "split": 3, "name": "shs"})
    opt = FusedAdam(groups, lr=0.0, eps=1e-15)
    # torch reference: dc / rest as separate tensors, as the reference model holds them
    tp = {k: params[k].clone().requires_grad_() for k in lrs}
    dc = params["shs"][:, :1].clone().requires_grad_()
    rest = params["shs"][:, 1:].clone().requires_grad_()
    topt = torch.optim.Adam([{"params": [tp[k]], "lr": lrs[k]} for k in lrs] +
                            [{"params": [dc], "lr": 0.0025}, {"params": [rest], "lr": 0.0025 / 20}], lr=0.0, eps=1e-15)
    for t in range(3):
        buf.flat.copy_(torch.randn(buf.flat.numel(), generator=gen).cuda() * NV)
        for k in lrs:
            tp[k].grad = buf.views[k].clone() / NV
        dc.grad = buf.views["shs"][:, :1].clone() / NV
        rest.grad = buf.views["shs"][:, 1:].clone() / NV
        topt.step()
        m2d = buf.views["means2D"].clone()
        opt.step(grads=buf.views, grad_scale=1.0 / NV, zero_grad=True)
        assert torch.equal(buf.views["means2D"], m2d)  # not a parameter: untouched
        for k in lrs:
            assert not buf.views[k].any()
            assert torch.all((params[k] - tp[k].detach()).abs() <= 2e-6 * (tp[k].detach().abs() + lrs[k])), (k, t)
        ref = torch.cat([dc.detach(), rest.detach()], 1)
        assert torch.all((params["shs"] - ref).abs() <= 2e-6 * (ref.abs() + 0.0025)), t


def test_adam_state_dict_round_trip_and_grad_lookup_errors():
    """torch.optim's checkpoint surface (the reference's capture() stores optimizer.state_dict(), mlp_flex.py:102)
    and the loud failures of name-keyed gradients."""
    from dmgs_b200.optim import FusedAdam
    gen = torch.Generator().manual_seed(4)
    mk = lambda: {k: torch.randn(1000, w, generator=torch.Generator().manual_seed(7)).cuda() for k, w in (("a", 3), ("b", 4))}
    pa, pb = mk(), mk()
    opt_a = FusedAdam([{"params": [pa[k]], "lr": 0.01, "name": k} for k in pa], lr=0.0, eps=1e-15)
    grads = [{k: torch.randn(v.shape, generator=gen).cuda() for k, v in pa.items()} for _ in range(4)]
    for g in grads[:2]:
        opt_a.step(grads={k: v.clone() for k, v in g.items()})
    sd = opt_a.state_dict()
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["params"] == [0] and float(sd["state"][1]["step"]) == 2
    # torch.optim.Adam accepts the same dict (same layout)
    tp = [pa["a"].clone().requires_grad_(), pa["b"].clone().requires_grad_()]
    topt = torch.optim.Adam([{"params": [tp[0]], "lr": 0.01}, {"params": [tp[1]], "lr": 0.01}], lr=0.0, eps=1e-15)
    topt.load_state_dict({"state": {i: {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in st.items()}
                                    for i, st in sd["state"].items()},
                          "param_groups": [{**topt.param_groups[i], "params": [i]} for i in range(2)]})
    opt_b = FusedAdam([{"params": [pb[k]], "lr": 0.5, "name": k} for k in pb], lr=0.0, eps=1e-15)
    for k in pb:
        pb[k].copy_(pa[k])
    opt_b.load_state_dict(sd)
    assert opt_b.param_groups[0]["lr"] == 0.01
    for g in grads[2:]:
        opt_a.step(grads={k: v.clone() for k, v in g.items()})
        opt_b.step(grads={k: v.clone() for k, v in g.items()})
        tp[0].grad, tp[1].grad = g["a"].clone(), g["b"].clone()
        topt.step()
    for k, t in zip(pa, tp):
        assert torch.equal(pa[k], pb[k])
        assert torch.all((pa[k] - t.detach()).abs() <= 2e-6 * (t.detach().abs() + 0.01))
    with pytest.raises(KeyError):
        opt_a.step(grads={"a": grads[0]["a"]})
    two = FusedAdam([{"params": [pa["a"], pa["b"]], "lr": 0.01, "name": "both"}])
    with pytest.raises(RuntimeError):
        two.step(grads={"both": grads[0]["a"]})
