"""Mesh-face -> Gaussian binding, fused on the device (libdmgs_raster.so: dmgs_bind_forward/backward).

Host-side mirror of the binding interface of the reference's stage-2/3 models:
  * ``bind_faces``        <- scene/gaussian_geo_model_mlp_flex.py:267-311 (renew_gaussian: frame,
                             barycentric means, affine cov3D_L) + :370-385 (get_covariance_dyn);
                             the COLMAP variant (…_mlp_flex_colmap.py:494-512) is ``scale_factor=None``.
  * ``bind_frame``        <- scene/gaussian_geo_model_finetune.py:414-421 (rot_t2w + means, all with grad)
  * ``stage3_scales_rotations`` / ``stage3_covariance`` <- …_finetune.py:446-482, :501-516
  * ``renew_gaussian``    <- the gs_info dict of mlp_flex.py:321-334 that render_dyn consumes.
Gradients follow the reference's autograd exactly: cov3D_L is a constant (built under no_grad,
mlp_flex.py:285), gradients reach ``verts`` through the face frame and the barycentric means and
reach ``scale_factor`` through Sigma.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .synthetic import barycentric_layout  # geo/mesh_utils.py:16-40 constants


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _BindFaces(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, faces, bc, g, rad_base, thin_z, adaptive, want_cov, want_rot):
        if verts.device.type != "cuda":
            raise RuntimeError("dmgs_b200 binding needs CUDA tensors; there is no CPU path")
        verts_c = verts.detach().float().contiguous()
        faces_c = faces.to(torch.int64).contiguous()
        bc_c = bc.detach().float().contiguous()
        g_c = None if g is None else g.detach().float().reshape(1).contiguous()
        F, k = int(faces_c.shape[0]), int(bc_c.shape[0])
        dev = verts.device
        xyz = torch.empty(F * k, 3, dtype=torch.float32, device=dev)
        cov6 = torch.empty(F * k, 6, dtype=torch.float32, device=dev) if want_cov else None
        rot = torch.empty(F, 3, 3, dtype=torch.float32, device=dev) if want_rot else None
        L.check(L.lib().dmgs_bind_forward(F, k, L.ptr(verts_c), L.ptr(faces_c), L.ptr(bc_c), float(rad_base),
                                          float(thin_z), L.ptr(g_c), int(adaptive), L.ptr(xyz), L.ptr(cov6),
                                          L.ptr(rot), _stream()), "dmgs_bind_forward")
        ctx.save_for_backward(verts_c, faces_c, bc_c, g_c if g_c is not None else torch.empty(0, device=dev))
        ctx.meta = (float(rad_base), float(thin_z), int(adaptive), g is not None, want_cov, want_rot)
        outs = [xyz]
        if want_cov:
            outs.append(cov6)
        if want_rot:
            outs.append(rot)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        verts_c, faces_c, bc_c, g_c = ctx.saved_tensors
        rad_base, thin_z, adaptive, has_g, want_cov, want_rot = ctx.meta
        it = iter(grads)
        fix = lambda t: None if t is None else t.float().contiguous()
        g_xyz = fix(next(it))
        g_cov = fix(next(it)) if want_cov else None
        g_rot = fix(next(it)) if want_rot else None
        F, k = int(faces_c.shape[0]), int(bc_c.shape[0])
        dverts = torch.zeros_like(verts_c)
        dg = torch.zeros(1, dtype=torch.float32, device=verts_c.device)
        L.check(L.lib().dmgs_bind_backward(F, k, L.ptr(verts_c), L.ptr(faces_c), L.ptr(bc_c), rad_base, thin_z,
                                           L.ptr(g_c if has_g else None), adaptive, L.ptr(g_xyz), L.ptr(g_cov),
                                           L.ptr(g_rot), L.ptr(dverts), L.ptr(dg), _stream()), "dmgs_bind_backward")
        return dverts, None, None, (dg if has_g else None), None, None, None, None, None


def bind_faces(verts, faces, bc, rad_base, thin_z, scale_factor=None, max_scale=2.0, adaptive_cov=True):
    """-> (gs_xyz [F*k,3], cov3D_precomp [F*k,6]), face-major Gaussian order (f*k + j)."""
    g = None if scale_factor is None else torch.tanh(scale_factor) * max_scale  # mlp_flex.py:377
    xyz, cov6 = _BindFaces.apply(verts, faces, bc, g, rad_base, thin_z, adaptive_cov, True, False)
    return xyz, cov6


def bind_frame(verts, faces, bc):
    """Stage-3 frame + means: -> (gs_xyz [F*k,3], rot_t2w [F,3,3]) (finetune.py:414-421)."""
    xyz, rot = _BindFaces.apply(verts, faces, bc, None, 0.0, 0.0, True, False, True)
    return xyz, rot


def renew_gaussian(verts, faces, bc, rad_base, spatial_lr_scale, scale_factor, features, max_scale=2.0,
                   adaptive_cov=True, active_sh_degree=3, max_sh_degree=3):
    """The gs_info dict of mlp_flex.py:321-334 (minus the FlexiCubes regularisers), for render_dyn."""
    xyz, cov = bind_faces(verts, faces, bc, rad_base, spatial_lr_scale * 1e-6, scale_factor, max_scale, adaptive_cov)
    N = xyz.shape[0]
    return {"xyz": xyz, "opacity": torch.full((N, 1), 0.9999, device=xyz.device), "covariance": cov,
            "features": features, "active_sh_degree": active_sh_degree, "max_sh_degree": max_sh_degree,
            "verts": verts, "faces": faces}


# ------------------------------------------------------------------------------ stage 3
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


def in_frustum(full_proj_transform, points):
    """Face-centroid frustum mask of finetune.py:33-48 (|ndc| < 1.05 and w > 0)."""
    p = points @ full_proj_transform[:3, :] + full_proj_transform[3:, :]
    w = p[:, 3:] + 1e-6
    ndc = p[:, :3] / w
    return (ndc.abs() < 1.05).all(dim=-1) & (w.squeeze(-1) > 0)


__all__ = ["bind_faces", "bind_frame", "renew_gaussian", "stage3_scales_rotations", "stage3_covariance",
           "stage3_rot_matrix", "matrix_to_quaternion", "in_frustum", "barycentric_layout"]
