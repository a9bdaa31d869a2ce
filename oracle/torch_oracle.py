"""fp64 autograd formulation of the splat math -- TEST INFRASTRUCTURE ONLY.

An independent, dense (every pixel x every Gaussian) PyTorch float64 statement of the same
forward model as oracle/splat_oracle.c.  Gradients come from autograd,
not from hand-derived formulas, so they are the gradient truth the C oracle's analytic backward
is validated against on tiny scenes (tests/test_oracle.py).  Two places deliberately reproduce
the upstream rasteriser's gradient conventions instead of the exact derivative (A.5/A.6):
``min(0.99, alpha)`` passes gradient through, and the frustum clamp of t.x/t.y is treated as a
constant when active.  Also used to restate utils/sh_utils.py:41-99 (eval_sh) in vector form.
"""
from __future__ import annotations

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def sh_basis(deg: int, d: torch.Tensor) -> torch.Tensor:
    """d: [P,3] unit directions -> basis [P,(deg+1)^2] with the signs of utils/sh_utils.py:75-99."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    b = [torch.full_like(x, C0)]
    if deg > 0:
        b += [-C1 * y, C1 * z, -C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
    if deg > 2:
        b += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
              C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
              C3[6] * x * (xx - 3 * yy)]
    return torch.stack(b, dim=1)


def eval_sh_colors(deg, shs_pmc, means3D, campos, act):
    """shs_pmc: [P,M,3]; returns rgb [P,3]. act 0: max(v+0.5,0); act 1: sigmoid."""
    d = means3D - campos[None]
    d = d / d.norm(dim=1, keepdim=True)
    b = sh_basis(deg, d)
    v = (b[:, :, None] * shs_pmc[:, : b.shape[1], :]).sum(1)
    return torch.clamp_min(v + 0.5, 0.0) if act == 0 else torch.sigmoid(v)


def quat_to_rot(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1)
    return R.view(-1, 3, 3)


def render(W, H, tanfovx, tanfovy, bg, view, proj, campos, means3D, opacities, scales=None, rotations=None,
           cov3D_precomp=None, shs=None, colors_precomp=None, sh_degree=3, sh_layout=0, sh_act=0,
           scale_modifier=1.0):
    """All tensors float64.  view/proj are the flat (transposed) matrices as the rasteriser gets them.
    Returns image [3,H,W] plus a dict of intermediates."""
    dt = means3D.dtype
    P = means3D.shape[0]
    Vt = view.reshape(4, 4).to(dt)  # = (W2C)^T
    PVt = proj.reshape(4, 4).to(dt)
    ones = torch.ones(P, 1, dtype=dt)
    ph = torch.cat([means3D, ones], 1)
    t = ph @ Vt  # [P,4] view space
    hom = ph @ PVt
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ppx, ppy = hom[:, 0] * pw, hom[:, 1] * pw
    tz = t[:, 2]
    vis = tz > 0.2
    if cov3D_precomp is None:
        R = quat_to_rot(rotations)
        S = torch.diag_embed(scale_modifier * scales)
        Mm = R @ S
        Sig = Mm @ Mm.transpose(1, 2)
    else:
        c = cov3D_precomp
        Sig = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], 1).view(-1, 3, 3)
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tzs = torch.where(vis, tz, torch.ones_like(tz))
    txtz, tytz = t[:, 0] / tzs, t[:, 1] / tzs
    cx = torch.where((txtz < -limx) | (txtz > limx), (txtz.clamp(-limx, limx) * tzs).detach(), t[:, 0])
    cy = torch.where((tytz < -limy) | (tytz > limy), (tytz.clamp(-limy, limy) * tzs).detach(), t[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tzs, zero, -(fx * cx) / (tzs * tzs), zero, fy / tzs, -(fy * cy) / (tzs * tzs)], 1).view(-1, 2, 3)
    Wr = Vt[:3, :3].transpose(0, 1)  # true world->view rotation
    Tm = J @ Wr[None]
    cov2 = Tm @ Sig @ Tm.transpose(1, 2)
    a, b, c_ = cov2[:, 0, 0] + 0.3, cov2[:, 0, 1], cov2[:, 1, 1] + 0.3
    det = a * c_ - b * b
    vis = vis & (det != 0)
    dets = torch.where(vis, det, torch.ones_like(det))
    conx, cony, conz = c_ / dets, -b / dets, a / dets
    mid = 0.5 * (a + c_)
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    rad = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    px, py = ((ppx + 1.0) * W - 1.0) * 0.5, ((ppy + 1.0) * H - 1.0) * 0.5
    gx, gy = (W + 15) // 16, (H + 15) // 16
    x0 = torch.clamp(((px - rad) / 16).detach(), 0, gx).to(torch.int64)
    x1 = torch.clamp(((px + rad + 15) / 16).detach(), 0, gx).to(torch.int64)
    y0 = torch.clamp(((py - rad) / 16).detach(), 0, gy).to(torch.int64)
    y1 = torch.clamp(((py + rad + 15) / 16).detach(), 0, gy).to(torch.int64)
    vis = vis & ((x1 - x0) * (y1 - y0) > 0)
    if colors_precomp is None:
        shs_pmc = shs if sh_layout == 0 else shs.transpose(1, 2)
        rgb = eval_sh_colors(sh_degree, shs_pmc, means3D, campos.to(dt), sh_act)
    else:
        rgb = colors_precomp
    # depth order (stable in index)
    order = torch.argsort(torch.where(vis, tz, torch.full_like(tz, float("inf"))).detach(), stable=True)
    order = order[: int(vis.sum())]
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pxf, pyf = xs.reshape(-1).to(dt), ys.reshape(-1).to(dt)
    txi, tyi = (xs.reshape(-1) // 16), (ys.reshape(-1) // 16)
    o = order
    in_rect = (txi[:, None] >= x0[o][None]) & (txi[:, None] < x1[o][None]) & \
              (tyi[:, None] >= y0[o][None]) & (tyi[:, None] < y1[o][None])
    dx = px[o][None] - pxf[:, None]
    dy = py[o][None] - pyf[:, None]
    power = -0.5 * (conx[o][None] * dx * dx + conz[o][None] * dy * dy) - cony[o][None] * dx * dy
    G = torch.exp(torch.clamp_max(power, 0.0))
    araw = opacities.reshape(-1)[o][None] * G
    alpha = araw - torch.clamp_min(araw - 0.99, 0.0).detach()  # min(0.99, .) with pass-through gradient
    contrib = in_rect & (power <= 0) & (alpha >= 1.0 / 255.0)
    am = torch.where(contrib, alpha, torch.zeros_like(alpha))
    one_m = 1.0 - am
    Tincl = torch.cumprod(one_m, dim=1)
    Tbefore = torch.cat([torch.ones_like(Tincl[:, :1]), Tincl[:, :-1]], 1)
    stop = contrib & (Tincl < 1e-4)
    dead = torch.cumsum(stop.to(torch.int64), dim=1) > 0  # this one and everything behind it
    live = contrib & ~dead
    w = torch.where(live, am * Tbefore, torch.zeros_like(am))
    Cimg = w @ rgb[o]
    Tfinal = torch.where(live, one_m, torch.ones_like(one_m)).prod(dim=1)
    img = Cimg + Tfinal[:, None] * bg.to(dt)[None]
    n_contrib = torch.where(live, torch.arange(1, o.numel() + 1)[None].expand_as(live), torch.zeros_like(live, dtype=torch.int64)).max(dim=1).values \
        if o.numel() > 0 else torch.zeros(H * W, dtype=torch.int64)
    inter = {"radii": torch.where(vis, rad, torch.zeros_like(rad)).to(torch.int32), "xy": torch.stack([px, py], 1),
             "conic": torch.stack([conx, cony, conz], 1), "rgb": rgb, "depth": tz, "order": o,
             "final_T": Tfinal.view(H, W), "n_contrib_in_visible_order": n_contrib.view(H, W), "vis": vis}
    return img.t().reshape(3, H, W), inter


# ------------------------------------------------------------------------------ stage-3 binding
# CPU restatement (torch ops, autograd gives the gradients) of
#   get_rot_matrix   /root/reference/scene/gaussian_geo_model_finetune.py:465-482
#   get_scaling      :446-453          get_covariance  :501-516
#   get_rotation     :456-463 = normalize(pytorch3d.transforms.matrix_to_quaternion(R)); pytorch3d 0.7.x is a
#                    dependency that is absent here: its published algorithm (candidate with the largest
#                    denominator, clamp 0.1) is restated in matrix_to_quaternion below.
# PINNED by tests/golden/binding_stage3.npz (R, scales, cov6 and the autograd gradients, generated by
# executing the reference lines); the quaternion conversion is checked through R(q) == R.
def matrix_to_quaternion(matrix: torch.Tensor) -> torch.Tensor:
    """Rotation matrices [N,3,3] -> quaternions (w,x,y,z) with non-negative... largest-denominator
    branch selection, the semantics of pytorch3d.transforms.matrix_to_quaternion that
    finetune.py:461 calls (restated; pytorch3d is not a dependency)."""
    m = matrix
    m00, m01, m02 = m[:, 0, 0], m[:, 0, 1], m[:, 0, 2]
    m10, m11, m12 = m[:, 1, 0], m[:, 1, 1], m[:, 1, 2]
    m20, m21, m22 = m[:, 2, 0], m[:, 2, 1], m[:, 2, 2]
    pos_sqrt = lambda x: torch.where(x > 0, torch.sqrt(torch.clamp_min(x, 1e-30)), torch.zeros_like(x))
    q_abs = torch.stack([pos_sqrt(1.0 + m00 + m11 + m22), pos_sqrt(1.0 + m00 - m11 - m22),
                         pos_sqrt(1.0 - m00 + m11 - m22), pos_sqrt(1.0 - m00 - m11 + m22)], dim=-1)
    cand = torch.stack([
        torch.stack([q_abs[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[:, 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[:, 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[:, 3] ** 2], dim=-1)], dim=-2)
    cand = cand / (2.0 * q_abs[:, :, None].clamp_min(0.1))
    best = q_abs.argmax(dim=-1)
    return cand[torch.arange(m.shape[0], device=m.device), best]


def stage3_rot_matrix(rot_t2w, rotation2d):
    """R = rot_t2w[f] @ [[a,-b,0],[b,a,0],[0,0,1]], (a,b) = normalize(_rotation) (finetune.py:465-482)."""
    F = rot_t2w.shape[0]
    c = torch.nn.functional.normalize(rotation2d, dim=1)
    a, b = c[:, 0], c[:, 1]
    z, o = torch.zeros_like(a), torch.ones_like(a)
    Rt = torch.stack([a, -b, z, b, a, z, z, z, o], dim=1).view(F, -1, 3, 3)
    return torch.matmul(rot_t2w.view(F, 1, 3, 3), Rt).view(-1, 3, 3)


def stage3_scales_rotations(rot_t2w, scaling2d, rotation2d, thin_z_scale):
    """-> (scales [P,3], rotations [P,4] unit (w,x,y,z)) as finetune.py:446-463 hands to render()."""
    s = torch.cat([torch.exp(scaling2d), torch.full((scaling2d.shape[0], 1), float(thin_z_scale),
                                                    device=scaling2d.device)], dim=1)
    q = matrix_to_quaternion(stage3_rot_matrix(rot_t2w, rotation2d))
    return s, torch.nn.functional.normalize(q)


def stage3_covariance(rot_t2w, scaling2d, rotation2d, thin_z_scale):
    """Sigma = (R S)(R S)^T stripped to 6 (finetune.py:501-516)."""
    s = torch.cat([torch.exp(scaling2d), torch.full((scaling2d.shape[0], 1), float(thin_z_scale),
                                                    device=scaling2d.device)], dim=1)
    Lm = stage3_rot_matrix(rot_t2w, rotation2d) * s[:, None, :]
    Sg = Lm @ Lm.transpose(1, 2)
    return torch.stack([Sg[:, 0, 0], Sg[:, 0, 1], Sg[:, 0, 2], Sg[:, 1, 1], Sg[:, 1, 2], Sg[:, 2, 2]], dim=1)


