"""Stage-2 colour field: hash-grid encoder + MLP, geo/texture.py:47-111.

PARITY STATUS: unpinned -- the encoder's arithmetic is tinycudann's, which is neither in /root/reference nor in this
image; oracle/texture_oracle.py restates its published definition.  What IS checked: the oracle's own consistency
(level table against the C library's, exact interpolation on dense levels, autograd against finite differences),
the MLP against torch's own Linear / ReLU (the reference's ops), and the CUDA kernels against the oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import texture_oracle as T


def _params(seed=0, channels=48, amp=0.05):
    grid, ws = T.init_params(seed, channels)
    return grid * (amp / 1e-4), ws  # a trained-looking table: the 1e-4 initialisation gives outputs ~ 0


def test_level_table_matches_the_library():
    from dmgs_b200 import _lib as L
    scales, res, offs = T.level_table()
    assert T.n_grid_params() == L.lib().dmgs_texture_grid_params() == 12599920
    assert res[0] == 16 and abs(float(scales[-1]) - 4095.0) < 0.5 and int(res[-1]) in (4096, 4097)
    sizes = np.diff(offs)
    assert (sizes <= 1 << 19).all() and (sizes % 8 == 0).all()
    assert (sizes[5:] == 1 << 19).all()  # hashed levels


def test_dense_level_interpolates_exactly():
    """On a dense level a table holding a linear function of the cell coordinates is reproduced exactly."""
    scales, res, offs = T.level_table()
    grid = torch.zeros(T.n_grid_params(), dtype=torch.float64)
    r = int(res[0])
    cells = torch.arange(r ** 3)
    cx, cy, cz = cells % r, (cells // r) % r, cells // (r * r)
    f0 = 0.1 * cx + 0.2 * cy - 0.05 * cz
    grid[2 * cells] = f0.double()
    grid[2 * cells + 1] = 1.0
    # (t <= 0.9: in the last cell the upper corner has index `resolution`, which the dense index wraps -- tinycudann's
    # resolution is ceil(scale) + 1 while positions reach scale + 0.5)
    t = torch.rand(64, 3, generator=torch.Generator().manual_seed(1), dtype=torch.float64) * 0.9
    enc = T.encode(t, grid, half=False)
    pos = t * float(scales[0]) + 0.5
    want = 0.1 * pos[:, 0] + 0.2 * pos[:, 1] - 0.05 * pos[:, 2]
    assert torch.allclose(enc[:, 0], want, atol=1e-12) and torch.allclose(enc[:, 1], torch.ones(64, dtype=torch.float64), atol=1e-12)


def test_oracle_gradients_match_finite_differences():
    grid, ws = _params(2, 8)
    grid, ws = grid.double(), [w.double() for w in ws]
    aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.5, 2.0]], dtype=torch.float64)
    xyz = (torch.rand(4, 3, generator=torch.Generator().manual_seed(3), dtype=torch.float64) * 1.6 - 0.8).requires_grad_()
    # element-wise for the positions (the trilinear weights), one random direction for the weight matrices
    f = lambda x: T.sample_noact(x, aabb, grid, ws, half=False)
    assert torch.autograd.gradcheck(f, (xyz,), eps=1e-7, atol=1e-6, rtol=1e-4)
    wr = [w.clone().requires_grad_() for w in ws]
    dirs = [torch.randn(w.shape, generator=torch.Generator().manual_seed(4 + i), dtype=torch.float64) for i, w in enumerate(ws)]
    proj = torch.randn(4, 8, generator=torch.Generator().manual_seed(9), dtype=torch.float64)
    loss = lambda w3: (T.sample_noact(xyz.detach(), aabb, grid, w3, half=False) * proj).sum()
    loss(wr).backward()
    analytic = sum((w.grad * d).sum() for w, d in zip(wr, dirs)).item()
    h = 1e-6
    numeric = (loss([w + h * d for w, d in zip(ws, dirs)]) - loss([w - h * d for w, d in zip(ws, dirs)])).item() / (2 * h)
    assert abs(analytic - numeric) <= 1e-6 * max(1.0, abs(numeric)), (analytic, numeric)


def test_mirror_has_the_reference_parameter_names():
    import dmgs_b200.texture as TX
    import inspect
    src = inspect.getsource(TX)
    assert "self.encoder = _HashGrid" in src and "self.net = _MLP" in src and "self.params = nn.Parameter" in src
    m = TX._MLP({"n_input_dims": 32, "n_output_dims": 48, "n_hidden_layers": 2, "n_neurons": 32})
    assert list(m.state_dict()) == ["net.0.weight", "net.2.weight", "net.4.weight"]  # -> net.net.{0,2,4}.weight in MLPTexture3D
    assert m.net[4].weight.shape == (48, 32) and m.net[0].bias is None


# ------------------------------------------------------------------------------------------------ GPU
def _module(channels=48, seed=0, amp=0.05):
    from dmgs_b200.texture import MLPTexture3D
    grid, ws = _params(seed, channels, amp)
    aabb = torch.tensor([[-1.0, -1.2, -0.9], [1.1, 1.0, 1.3]])
    tex = MLPTexture3D(aabb.cuda(), channels=channels)
    with torch.no_grad():
        tex.encoder.params.copy_(grid.cuda())
        for m, w in zip(tex.net.weights(), ws):
            m.copy_(w.cuda())
    return tex, aabb, grid, ws


@pytest.mark.gpu
@pytest.mark.parametrize("channels", [48, 4])
def test_texture_forward_matches_oracle(channels):
    tex, aabb, grid, ws = _module(channels)
    assert list(tex.state_dict()) == ["encoder.params", "net.net.0.weight", "net.net.2.weight", "net.net.4.weight"]
    g = torch.Generator().manual_seed(5)
    xyz = torch.rand(3000, 3, generator=g) * 2.6 - 1.3  # some points outside the box: clamped
    out = tex.sample_noact(xyz.cuda().view(30, 100, 3))
    assert out.shape == (30, 100, channels)
    ref = T.sample_noact(xyz.double(), aabb.double(), grid.double(), [w.double() for w in ws], half=True, f32_coords=True)
    err = (out.view(-1, channels).cpu().double() - ref).abs()
    scale = ref.abs().max().item()
    # the encoder output is rounded to fp16 on both sides; a feature that sits on a rounding boundary may round the
    # other way (1 half ulp = 5e-4 relative of that feature) -- the bulk must agree to fp32 accuracy
    assert err.max().item() <= 2e-3 * scale, (err.max().item(), scale)
    # (fp32 trilinear weights + three fp32 layers of 32 terms each: a few 1e-6 of the output scale)
    assert err.median().item() <= 5e-6 * scale and (err > 1e-4 * scale).float().mean().item() < 0.02
    assert tex.sample_noact(torch.zeros(0, 3, device="cuda")).shape == (0, channels)


@pytest.mark.gpu
def test_texture_backward_matches_oracle_autograd():
    channels = 48
    tex, aabb, grid, ws = _module(channels, seed=7)
    g = torch.Generator().manual_seed(11)
    xyz = torch.rand(1600, 3, generator=g) * 2.4 - 1.2
    # ReLU has a kink: a unit whose pre-activation is within rounding of zero is "on" in one precision and "off" in
    # the other, and the whole adjoint chain of that point changes.  Points with such a unit are left out (a handful
    # in 1600); everything else must agree.
    with torch.no_grad():
        enc = T.encode(torch.clamp(((xyz.double() - aabb[0].double()) / (aabb[1] - aabb[0]).double()).float().double(), 0, 1),
                       grid.double(), half=True, f32_coords=True)
        p1 = enc @ ws[0].double().T
        p2 = torch.relu(p1) @ ws[1].double().T
        keep = (p1.abs().min(1).values > 1e-5) & (p2.abs().min(1).values > 1e-5)
    assert keep.float().mean() > 0.95
    xyz = xyz[keep]
    N = xyz.shape[0]
    dL = torch.randn(N, channels, generator=g)
    x = xyz.cuda().requires_grad_()
    out = tex.sample_noact(x)
    (out * dL.cuda()).sum().backward()
    # oracle: fp64 autograd with straight-through fp16 roundings
    gd, wd = grid.double().requires_grad_(), [w.double().requires_grad_() for w in ws]
    xd = xyz.double().requires_grad_()
    ref = T.sample_noact(xd, aabb.double(), gd, wd, half=True, f32_coords=True)
    (ref * dL.double()).sum().backward()

    def close(got, want, name, rtol=2e-3):
        got, want = got.detach().cpu().double(), want.detach()
        tol = rtol * want.abs().max().item()
        err = (got - want).abs().max().item()
        assert err <= tol, f"{name}: max err {err:.3e} > {tol:.3e}"

    for m, w, nm in zip(tex.net.weights(), wd, ("W0", "W1", "W2")):
        close(m.grad, w.grad, nm)
    # the reference's hooks hand the optimiser 128 x the true gradient of the encoder parameters
    close(tex.encoder.params.grad, gd.grad * T.GRAD_SCALE, "encoder.params")
    nz = gd.grad != 0
    assert torch.equal(tex.encoder.params.grad.cpu() != 0, nz) or (tex.encoder.params.grad.cpu() != 0)[nz].float().mean() > 0.999
    close(x.grad, xd.grad, "xyz")
    outside = ((xyz < aabb[0]) | (xyz > aabb[1]))
    assert outside.any() and (x.grad.cpu()[outside] == 0).all()  # clamped coordinates pass no gradient


@pytest.mark.gpu
def test_texture_feeds_the_fused_stage2_render():
    """features = sample_noact(gs_xyz).view(N, 3, 16) (mlp_flex.py:313) into the sigmoid-SH rasteriser, end to end:
    the loss moves when the texture parameters take a gradient step."""
    from dmgs_b200 import GaussianRasterizer, synthetic as S
    from gpu_util import settings_for
    tex, aabb, grid, ws = _module(48, seed=3, amp=0.5)
    cl = S.random_cloud(4000, seed=2, extent=0.8, log_scale_mean=math.log(0.05))
    cam = S.nerf_synthetic_camera(0, 160, 120)
    ras = GaussianRasterizer(settings_for(cam, (0, 0, 0)), sh_activation="sigmoid", sh_layout="P3M")
    gt = torch.rand(3, 120, 160, generator=torch.Generator().manual_seed(1)).cuda()
    xyz = cl["means3D"].cuda()
    opt = torch.optim.Adam(tex.parameters(), lr=1e-2)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        feats = tex.sample_noact(xyz).view(-1, 3, 16)
        img, _ = ras(means3D=xyz, means2D=torch.zeros_like(xyz), shs=feats, opacities=cl["opacities"].cuda(),
                     scales=cl["scales"].cuda(), rotations=cl["rotations"].cuda())
        loss = (img - gt).abs().mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses
