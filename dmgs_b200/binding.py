"""Mesh-face -> Gaussian binding, fused on the device (libdmgs_raster.so: dmgs_bind_forward/backward).

Host-side mirror of the binding interface of the reference's stage-2/3 models:
  * ``bind_faces``        <- scene/gaussian_geo_model_mlp_flex.py:267-311 (renew_gaussian: frame,
                             barycentric means, affine cov3D_L) + :370-385 (get_covariance_dyn);
                             the COLMAP variant (…_mlp_flex_colmap.py:494-512) is ``scale_factor=None``.
  * ``bind_frame``        <- scene/gaussian_geo_model_finetune.py:414-421 (rot_t2w + means, all with grad)
  * ``stage3_scales_rotations`` / ``stage3_covariance`` <- …_finetune.py:446-482, :501-516 (fused:
                             dmgs_stage3_forward/backward)
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


def _stream(dev=None):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _BindFaces(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, faces, bc, g, rad_base, thin_z, adaptive, want_cov, want_rot):
        if verts.device.type != "cuda":
            raise RuntimeError("dmgs_b200 binding needs CUDA tensors; there is no CPU path")
        if bc.dim() == 3 and bc.shape[0] == 1:  # the reference also uses bc_coords.view(1, -1, 3) (mlp_flex.py:281)
            bc = bc[0]
        if bc.dim() != 2 or bc.shape[1] != 3 or bc.shape[0] < 1:
            raise ValueError(f"bc: expected the [k,3] barycentric table (geo/mesh_utils.py:16-40), got {tuple(bc.shape)}")
        if faces.dim() != 2 or faces.shape[1] != 3:
            raise ValueError(f"faces: expected [F,3], got {tuple(faces.shape)}")
        if verts.dim() != 2 or verts.shape[1] != 3:
            raise ValueError(f"verts: expected [V,3], got {tuple(verts.shape)}")
        if faces.device != verts.device or bc.device != verts.device:
            raise ValueError("verts, faces and bc must live on the same CUDA device")
        verts_c = verts.detach().float().contiguous()
        faces_c = faces.to(torch.int64).contiguous()
        bc_c = bc.detach().float().contiguous()
        g_c = None if g is None else g.detach().float().reshape(1).contiguous()
        F, k = int(faces_c.shape[0]), int(bc_c.shape[0])
        dev = verts.device
        xyz = torch.empty(F * k, 3, dtype=torch.float32, device=dev)
        cov6 = torch.empty(F * k, 6, dtype=torch.float32, device=dev) if want_cov else None
        rot = torch.empty(F, 3, 3, dtype=torch.float32, device=dev) if want_rot else None
        with torch.cuda.device(dev):
            L.check(L.lib().dmgs_bind_forward(F, k, L.ptr(verts_c), L.ptr(faces_c), L.ptr(bc_c), float(rad_base),
                                              float(thin_z), L.ptr(g_c), int(adaptive), L.ptr(xyz), L.ptr(cov6),
                                              L.ptr(rot), _stream(dev)), "dmgs_bind_forward")
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
        with torch.cuda.device(verts_c.device):
            L.check(L.lib().dmgs_bind_backward(F, k, L.ptr(verts_c), L.ptr(faces_c), L.ptr(bc_c), rad_base, thin_z,
                                               L.ptr(g_c if has_g else None), adaptive, L.ptr(g_xyz), L.ptr(g_cov),
                                               L.ptr(g_rot), L.ptr(dverts), L.ptr(dg), _stream(verts_c.device)),
                    "dmgs_bind_backward")
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


def face_normals(vertices, faces, unit=False):
    """geo/mesh_utils.py:43-57 with unit=True (the only form on the path: mlp_flex.py:268, finetune.py:414):
    normalize((v1 - v0) x (v2 - v0)), taken from the fused frame kernel (third column of rot_t2w), with grad."""
    if not unit:
        raise NotImplementedError("only unit=True is on the binding path (mlp_flex.py:268, finetune.py:414)")
    bc = torch.full((1, 3), 1.0 / 3.0, dtype=torch.float32, device=vertices.device)
    _, rot = _BindFaces.apply(vertices, faces, bc, None, 0.0, 0.0, True, False, True)
    return rot[:, :, 2]


def renew_gaussian(verts, faces, bc, rad_base, spatial_lr_scale, scale_factor, features, max_scale=2.0,
                   adaptive_cov=True, active_sh_degree=3, max_sh_degree=3):
    """The gs_info dict of mlp_flex.py:321-334 (minus the FlexiCubes regularisers), for render_dyn."""
    xyz, cov = bind_faces(verts, faces, bc, rad_base, spatial_lr_scale * 1e-6, scale_factor, max_scale, adaptive_cov)
    N = xyz.shape[0]
    return {"xyz": xyz, "opacity": torch.full((N, 1), 0.9999, device=xyz.device), "covariance": cov,
            "features": features, "active_sh_degree": active_sh_degree, "max_sh_degree": max_sh_degree,
            "verts": verts, "faces": faces}


# ------------------------------------------------------------------------------ binding fused into the rasteriser
class _RasterizeMesh(torch.autograd.Function):
    """verts, g, features | colors, opacities, means2D -> image, radii (+ detached means): the stage-2 chain
    renew_gaussian -> get_covariance_dyn -> render_dyn (mlp_flex.py:267-311, :370-385; gaussian_renderer/__init__.py:
    104-193) with the binding inside the per-Gaussian kernels (dmgs_preprocess_forward_bound / _backward_bound)."""

    @staticmethod
    def forward(ctx, verts, g, shs, colors_precomp, opacities, means2D, faces, bc, rad_base, thin_z, adaptive, settings,
                sh_layout, sh_activation, want_xyz, holder):
        from . import rasterizer as RZ
        if verts.device.type != "cuda":
            raise RuntimeError("dmgs_b200 binding needs CUDA tensors; there is no CPU path")
        if bc.dim() == 3 and bc.shape[0] == 1:
            bc = bc[0]
        mesh = RZ.BoundMesh(verts.detach().float().contiguous(), faces.to(torch.int64).contiguous(),
                            bc.detach().float().contiguous(), float(rad_base), float(thin_z),
                            None if g is None else g.detach().float().reshape(1).contiguous(), bool(adaptive))
        P = int(mesh.faces.shape[0]) * int(mesh.bc.shape[0])
        sh_c, col_c = RZ._f32c(shs), RZ._f32c(colors_precomp)
        op_c = RZ._f32c(opacities)
        xyz = torch.empty(P, 3, dtype=torch.float32, device=verts.device) if want_xyz else None
        color, radii, state = RZ.rasterize_forward(settings, None, op_c, sh_c, col_c, None, None, None, sh_layout,
                                                   sh_activation, bound=mesh, xyz_out=xyz)
        ctx.state, ctx.mesh, ctx.has_col, ctx.has_g = state, mesh, col_c is not None, g is not None
        ctx.save_for_backward(sh_c)
        if holder is not None:
            holder.last = state
        ctx.mark_non_differentiable(radii)
        if want_xyz:
            ctx.mark_non_differentiable(xyz)
            return color, radii, xyz
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, *_unused):
        from . import rasterizer as RZ
        (sh_c,) = ctx.saved_tensors
        mesh = ctx.mesh
        dverts = torch.zeros_like(mesh.verts)
        dg = torch.zeros(1, dtype=torch.float32, device=dverts.device) if ctx.has_g else None
        g_means2D, g_shs, g_col, g_op = RZ.rasterize_backward_bound(ctx.state, grad_color.contiguous().float(), mesh, sh_c,
                                                                    ctx.has_col, dverts, dg)
        return (dverts, dg, g_shs, g_col, g_op, g_means2D) + (None,) * 10


class MeshRasterizer:
    """Holder mirroring GaussianRasterizer's `.last` for the fused mesh path."""

    last = None


def rasterize_mesh(raster_settings, verts, faces, bc, rad_base, thin_z, opacities, features=None, colors_precomp=None,
                   scale_factor=None, max_scale=2.0, adaptive_cov=True, means2D=None, sh_activation="sigmoid",
                   sh_layout="P3M", return_xyz=False, holder=None):
    """Stage-2 render straight from the mesh: -> (image [3,H,W], radii [F*k]) (+ detached gs_xyz [F*k,3] when
    return_xyz, for the texture MLP).  Same values and gradients as bind_faces(...) followed by
    GaussianRasterizer(settings, sh_activation, sh_layout)(means3D=gs_xyz, cov3D_precomp=cov, shs=features, ...),
    without the xyz / cov3D_precomp / dL/dxyz / dL/dcov tensors.  means2D ([F*k,3] zeros with requires_grad, the
    reference's screenspace_points) receives the viewspace gradient."""
    if (features is None) == (colors_precomp is None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    g = None if scale_factor is None else torch.tanh(scale_factor) * max_scale  # mlp_flex.py:377
    P = int(faces.shape[0]) * int(bc.shape[-2])
    if means2D is None:
        means2D = torch.zeros(P, 3, dtype=torch.float32, device=verts.device)
    act = {"clamp": 0, "sigmoid": 1}[sh_activation]
    lay = {"PM3": 0, "P3M": 1}[sh_layout]
    return _RasterizeMesh.apply(verts, g, features, colors_precomp, opacities, means2D, faces, bc, rad_base, thin_z,
                                adaptive_cov, raster_settings, lay, act, return_xyz, holder)


# ------------------------------------------------------------------------------ stage 3
class _Stage3(torch.autograd.Function):
    """(rot_t2w [F,3,3], scaling2d [P,2], rotation2d [P,2]) -> scales / quaternions / cov6 on the device
    (libdmgs_raster.so: dmgs_stage3_forward / dmgs_stage3_backward)."""

    @staticmethod
    def forward(ctx, rot_t2w, scaling2d, rotation2d, thin_z, want):
        if rot_t2w.device.type != "cuda":
            raise RuntimeError("dmgs_b200 binding needs CUDA tensors; there is no CPU path")
        rot = rot_t2w.detach().float().contiguous()
        s2 = scaling2d.detach().float().contiguous()
        r2 = rotation2d.detach().float().contiguous()
        F, P = int(rot.shape[0]), int(s2.shape[0])
        if F == 0 or P % max(F, 1) != 0 or r2.shape[0] != P:
            if P != 0 or F != 0:
                raise ValueError(f"stage 3: {P} Gaussians over {F} faces")
        k = P // F if F else 1
        dev = rot.device
        scales = torch.empty(P, 3, dtype=torch.float32, device=dev) if "scales" in want else None
        quats = torch.empty(P, 4, dtype=torch.float32, device=dev) if "quats" in want else None
        cov6 = torch.empty(P, 6, dtype=torch.float32, device=dev) if "cov6" in want else None
        with torch.cuda.device(dev):
            L.check(L.lib().dmgs_stage3_forward(F, k, L.ptr(rot), L.ptr(r2), L.ptr(s2), float(thin_z), L.ptr(scales),
                                                L.ptr(quats), L.ptr(cov6), _stream(dev)), "dmgs_stage3_forward")
        ctx.save_for_backward(rot, s2, r2)
        ctx.meta = (F, k, float(thin_z), want)
        return tuple(t for t in (scales, quats, cov6) if t is not None)

    @staticmethod
    def backward(ctx, *grads):
        rot, s2, r2 = ctx.saved_tensors
        F, k, thin_z, want = ctx.meta
        it = iter(grads)
        fix = lambda t: None if t is None else t.float().contiguous()
        g_s = fix(next(it)) if "scales" in want else None
        g_q = fix(next(it)) if "quats" in want else None
        g_c = fix(next(it)) if "cov6" in want else None
        d_rot, d_r2, d_s2 = torch.empty_like(rot), torch.empty_like(r2), torch.empty_like(s2)
        with torch.cuda.device(rot.device):
            L.check(L.lib().dmgs_stage3_backward(F, k, L.ptr(rot), L.ptr(r2), L.ptr(s2), thin_z, L.ptr(g_s), L.ptr(g_q),
                                                 L.ptr(g_c), L.ptr(d_rot), L.ptr(d_r2), L.ptr(d_s2), _stream(rot.device)),
                    "dmgs_stage3_backward")
        return d_rot, d_s2, d_r2, None, None


def stage3_scales_rotations(rot_t2w, scaling2d, rotation2d, thin_z_scale):
    """-> (scales [P,3], rotations [P,4] unit (w,x,y,z)) as finetune.py:446-463 hands to render():
    get_scaling, and get_rotation = normalize(matrix_to_quaternion(get_rot_matrix()))."""
    return _Stage3.apply(rot_t2w, scaling2d, rotation2d, thin_z_scale, ("scales", "quats"))


def stage3_covariance(rot_t2w, scaling2d, rotation2d, thin_z_scale):
    """Sigma = (R S)(R S)^T stripped to 6 (finetune.py:501-516)."""
    return _Stage3.apply(rot_t2w, scaling2d, rotation2d, thin_z_scale, ("cov6",))[0]


def in_frustum(full_proj_transform, points):
    """Face-centroid frustum mask of finetune.py:33-48 (|ndc| < 1.05 and w > 0); see dmgs_b200.frustum."""
    from .frustum import in_frustum as _f
    return _f(full_proj_transform, points)


__all__ = ["bind_faces", "bind_frame", "rasterize_mesh", "MeshRasterizer", "face_normals", "renew_gaussian", "stage3_scales_rotations", "stage3_covariance", "in_frustum",
           "barycentric_layout"]
