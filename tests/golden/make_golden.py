"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN PYTHON LINES on the CPU.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
The model classes cannot be imported (missing plyfile/trimesh/tinycudann/pytorch3d/...), and
utils/sh_utils.py + utils/general_utils.py hard-code CUDA, so this script slices the cited
source line ranges out of the reference files, dedents them and exec()s them unchanged, with
`device` bound to 'cpu' and torch.zeros/full/eye patched to ignore device="cuda".
Nothing is copied into the repo except the resulting numbers.
"""
import math
import os
import sys
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(OUT, "..", ".."))
sys.path.insert(0, REF)

_orig = {n: getattr(torch, n) for n in ("zeros", "full", "eye", "tensor")}


def _cpu(fn):
    def w(*a, **k):
        if "device" in k:
            k["device"] = "cpu"
        return fn(*a, **k)
    return w


for n, f in _orig.items():
    setattr(torch, n, _cpu(f))


def ref_lines(rel, lo, hi):
    with open(os.path.join(REF, rel)) as fh:
        lines = fh.readlines()
    return textwrap.dedent("".join(lines[lo - 1:hi]))


def main():
    from dmgs_b200 import synthetic as S
    import geo.mesh_utils as mesh_utils  # reference module, imports on CPU

    gen = torch.Generator().manual_seed(7)

    # ---------------- barycentric layouts: geo/mesh_utils.py:16-40
    bary = {}
    for k in (1, 3, 6):
        bc, r = mesh_utils.generate_barycentric_v2(k)
        bary[f"bc{k}"] = bc.numpy()
        bary[f"rad{k}"] = np.float64(r)
    np.savez(os.path.join(OUT, "barycentric.npz"), **bary)

    # ---------------- general_utils: strip_symmetric / build_rotation (utils/general_utils.py:64-110)
    gu = {"torch": torch}
    exec(ref_lines("utils/general_utils.py", 64, 110), gu)

    # ---------------- eval_sh (utils/sh_utils.py:25-110; the CUDA warm-up at :113-116 is not executed)
    shns = {"torch": torch}
    exec(ref_lines("utils/sh_utils.py", 23, 110).replace("@torch.jit.script", ""), shns)
    eval_sh = shns["eval_sh"]
    Psh = 64
    feats = torch.randn(Psh, 3, 16, generator=gen)
    xyz = torch.randn(Psh, 3, generator=gen) * 1.5
    campos = torch.tensor([0.3, -2.0, 1.1])
    d = xyz - campos[None]
    d = d / d.norm(dim=1, keepdim=True)  # gaussian_renderer/__init__.py:75-76
    sh_out = {"features": feats.numpy(), "xyz": xyz.numpy(), "campos": campos.numpy()}
    for deg in range(4):
        v = eval_sh(deg, feats, d)
        sh_out[f"raw{deg}"] = v.numpy()
        sh_out[f"sigmoid{deg}"] = torch.sigmoid(v).numpy()               # :78 / :170
        sh_out[f"clamp{deg}"] = torch.clamp_min(v + 0.5, 0.0).numpy()    # :79 / :171 (rasteriser variant)
    sh_out["zero"] = eval_sh(3, torch.zeros(4, 3, 16), torch.zeros(4, 3)).numpy()  # proto&effi_stage2.ipynb cell 6
    np.savez(os.path.join(OUT, "eval_sh.npz"), **sh_out)

    # ---------------- stage-2 binding: mlp_flex.py:267-310 + get_covariance_dyn :370-385
    bind_src = ref_lines("scene/gaussian_geo_model_mlp_flex.py", 267, 310)
    cov_src = ref_lines("scene/gaussian_geo_model_mlp_flex.py", 370, 385)
    out = {}
    for k in (1, 3, 6):
        for adaptive in (True, False):
            verts, faces = S.jittered_sphere_mesh(80, seed=k, jitter=0.15)
            verts = verts.clone().requires_grad_(True)
            bc, rad = mesh_utils.generate_barycentric_v2(k)
            me = types.SimpleNamespace(bc_coords=bc, adaptive_cov=adaptive, rad_base=rad, spatial_lr_scale=4.43,
                                       scale_factor=torch.nn.Parameter(torch.tensor([0.4 + 0.1 * k])), max_scale=2)
            ns = {"torch": torch, "mesh_utils": mesh_utils, "self": me, "verts": verts, "faces": faces,
                  "device": "cpu", "strip_symmetric": gu["strip_symmetric"]}
            exec(bind_src, ns)
            exec(cov_src, ns)
            cov = ns["get_covariance_dyn"](me, ns["rot_t2w"], ns["cov3D_L"])
            P = faces.shape[0] * k
            gx = torch.randn(P, 3, generator=gen)
            gc = torch.randn(P, 6, generator=gen)
            ((ns["gs_xyz"] * gx).sum() + (cov * gc).sum() * 1e3).backward()
            tag = f"k{k}_{'adp' if adaptive else 'iso'}"
            out.update({f"{tag}_verts": verts.detach().numpy(), f"{tag}_faces": faces.numpy(),
                        f"{tag}_scale_factor": me.scale_factor.detach().numpy(),
                        f"{tag}_xyz": ns["gs_xyz"].detach().numpy(), f"{tag}_rot_t2w": ns["rot_t2w"].detach().numpy(),
                        f"{tag}_cov3D_L": ns["cov3D_L"].numpy(), f"{tag}_cov6": cov.detach().numpy(),
                        f"{tag}_gxyz": gx.numpy(), f"{tag}_gcov": (gc * 1e3).numpy(),
                        f"{tag}_dverts": verts.grad.numpy(), f"{tag}_dscale_factor": me.scale_factor.grad.numpy()})
    # affine_verify.ipynb cell 0-1 triangle (example B)
    verts = torch.tensor([[0.0, 0.0, 0.0], [0.0038, 0.0, 0.0], [0.0011, 0.0035, 0.0]])
    faces = torch.tensor([[0, 1, 2]])
    bc, rad = mesh_utils.generate_barycentric_v2(6)
    me = types.SimpleNamespace(bc_coords=bc, adaptive_cov=True, rad_base=rad, spatial_lr_scale=4.43)
    ns = {"torch": torch, "mesh_utils": mesh_utils, "self": me, "verts": verts, "faces": faces, "device": "cpu"}
    exec(bind_src, ns)
    out["affineB_cov3D_L"] = ns["cov3D_L"].numpy()
    np.savez(os.path.join(OUT, "binding_stage2.npz"), **out)

    # ---------------- stage-3 binding: finetune.py:414-421 (frame, means), :465-482 (get_rot_matrix),
    # :501-516 (get_covariance), :33-48 (in_frustum)
    fr_src = ref_lines("scene/gaussian_geo_model_finetune.py", 414, 421)
    rot_src = ref_lines("scene/gaussian_geo_model_finetune.py", 465, 482)
    cov3_src = ref_lines("scene/gaussian_geo_model_finetune.py", 501, 516)
    fru_src = ref_lines("scene/gaussian_geo_model_finetune.py", 34, 48).replace("@torch.jit.script", "")
    k = 3
    verts, faces = S.jittered_sphere_mesh(80, seed=11, jitter=0.1)
    verts = verts.clone().requires_grad_(True)
    bc, rad = mesh_utils.generate_barycentric_v2(k)
    P = faces.shape[0] * k
    me = types.SimpleNamespace(bc_coords=bc, verts=verts, faces=faces, gs_mask=None,
                               _rotation=torch.randn(P, 2, generator=gen).requires_grad_(True),
                               _scaling=(torch.randn(P, 2, generator=gen) * 0.3 - 3.0).requires_grad_(True),
                               thin_z_scale=4.43e-6, scaling_activation=torch.exp)
    ns = {"torch": torch, "mesh_utils": mesh_utils, "self": me, "verts": verts, "faces": faces,
          "strip_symmetric": gu["strip_symmetric"]}
    exec(fr_src, ns)
    exec(rot_src, ns)
    exec(cov3_src, ns)
    me.get_rot_matrix = lambda: ns["get_rot_matrix"](me)
    Rg = me.get_rot_matrix()
    s3 = torch.cat([torch.exp(me._scaling), torch.full((P, 1), me.thin_z_scale)], dim=1)  # :446-453
    me.get_scaling = s3
    cov3 = ns["get_covariance"](me)
    gx = torch.randn(P, 3, generator=gen)
    gc = torch.randn(P, 6, generator=gen) * 1e3
    ((me._xyz * gx).sum() + (cov3 * gc).sum()).backward()
    cam = S.look_at_camera([2.5, 0.5, 0.8], 200, 150, fovx=0.5)
    exec(fru_src, ns)
    centroids = verts.detach()[faces].mean(dim=1)
    fmask = ns["in_frustum"](cam.full_proj_transform, centroids)
    np.savez(os.path.join(OUT, "binding_stage3.npz"), verts=verts.detach().numpy(), faces=faces.numpy(),
             rotation2d=me._rotation.detach().numpy(), scaling2d=me._scaling.detach().numpy(),
             thin_z=np.float32(me.thin_z_scale), xyz=me._xyz.detach().numpy(), rot_t2w=me.rot_t2w.detach().numpy(),
             R=Rg.detach().numpy(), scales3=s3.detach().numpy(), cov6=cov3.detach().numpy(), gxyz=gx.numpy(),
             gcov=gc.numpy(), dverts=verts.grad.numpy(), drotation2d=me._rotation.grad.numpy(),
             dscaling2d=me._scaling.grad.numpy(), full_proj=cam.full_proj_transform.numpy(),
             frustum_mask=fmask.numpy())

    # ---------------- camera conventions: utils/graphics_utils.py:38-71 + scene/cameras.py:54-57
    from utils.graphics_utils import getWorld2View2, getProjectionMatrix
    Rm = np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]])
    Tv = np.array([0.1, -0.2, 4.0])
    fovx, fovy = 0.69, 0.52
    wv = torch.tensor(getWorld2View2(Rm, Tv, np.array([0.0, 0.0, 0.0]), 1.0)).transpose(0, 1)
    pm = getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fovx, fovY=fovy).transpose(0, 1)
    full = (wv.unsqueeze(0).bmm(pm.unsqueeze(0))).squeeze(0)
    np.savez(os.path.join(OUT, "camera.npz"), R=Rm, T=Tv, fovx=fovx, fovy=fovy, world_view=wv.numpy(),
             projection=pm.numpy(), full_proj=full.numpy(), center=wv.inverse()[3, :3].numpy())

    # ---------------- covariance / quaternion layout: build_scaling_rotation + strip_symmetric
    q = torch.randn(32, 4, generator=gen)
    s = torch.rand(32, 3, generator=gen) * 0.1 + 0.01
    L = gu["build_scaling_rotation"](s, q)  # normalises q internally (general_utils.py:79-81)
    cov = gu["strip_symmetric"](L @ L.transpose(1, 2))
    np.savez(os.path.join(OUT, "cov_layout.npz"), q=q.numpy(), s=s.numpy(), R=gu["build_rotation"](q).numpy(),
             cov6=cov.numpy())
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
