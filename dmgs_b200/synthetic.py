"""Seeded synthetic inputs shared by the oracle, the CUDA path, the tests and bench.py.

Everything is generated on the CPU with ``torch.Generator().manual_seed(seed)`` so the oracle
and the GPU see identical bits.  Camera conventions follow the reference:
``getWorld2View2`` / ``getProjectionMatrix`` (utils/graphics_utils.py:38-71), the transposes and
``camera_center`` of ``Camera.__init__`` (scene/cameras.py:54-57) and the OpenGL->COLMAP flip
of ``readCamerasFromTransforms`` (scene/dataset_readers.py:192-199).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

NERF_SYNTHETIC_FOVX = 0.6911112070083618  # camera_angle_x of the NeRF-synthetic scenes
NERF_SYNTHETIC_RADIUS = 4.0311289


@dataclass
class SynthCamera:
    """Duck-typed stand-in for scene.cameras.Camera / MiniCam (only the fields render() reads)."""

    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor  # [4,4] = (W2C)^T
    full_proj_transform: torch.Tensor  # [4,4] = (W2C)^T P^T
    camera_center: torch.Tensor  # [3]
    znear: float = 0.01
    zfar: float = 100.0

    def to(self, device):
        return SynthCamera(self.image_width, self.image_height, self.FoVx, self.FoVy,
                           self.world_view_transform.to(device), self.full_proj_transform.to(device),
                           self.camera_center.to(device), self.znear, self.zfar)


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    """utils/graphics_utils.py:51-71."""
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    bottom, left = -top, -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def camera_from_c2w(c2w_opengl: np.ndarray, width: int, height: int, fovx: float, fovy: float,
                    znear: float = 0.01, zfar: float = 100.0) -> SynthCamera:
    c2w = np.array(c2w_opengl, dtype=np.float64)
    c2w[:3, 1:3] *= -1  # OpenGL (Y up, Z back) -> COLMAP (Y down, Z forward)
    w2c = np.linalg.inv(c2w)
    R = np.transpose(w2c[:3, :3])
    T = w2c[:3, 3]
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    wv = torch.tensor(np.float32(Rt)).transpose(0, 1).contiguous()
    proj = projection_matrix(znear, zfar, fovx, fovy).transpose(0, 1)
    full = (wv.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    center = wv.inverse()[3, :3].contiguous()
    return SynthCamera(width, height, fovx, fovy, wv, full, center, znear, zfar)


def look_at_camera(eye, width: int, height: int, fovx: float, fovy: float | None = None,
                   target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)) -> SynthCamera:
    eye = np.asarray(eye, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    back = eye - target
    back /= np.linalg.norm(back)
    right = np.cross(np.asarray(up, dtype=np.float64), back)
    right /= np.linalg.norm(right)
    upv = np.cross(back, right)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, upv, back, eye
    if fovy is None:
        focal = width / (2 * math.tan(fovx / 2))
        fovy = 2 * math.atan(height / (2 * focal))
    return camera_from_c2w(c2w, width, height, fovx, fovy)


def nerf_synthetic_camera(seed: int = 0, width: int = 800, height: int = 800,
                          radius: float = NERF_SYNTHETIC_RADIUS) -> SynthCamera:
    g = torch.Generator().manual_seed(1000 + seed)
    u = torch.rand(2, generator=g).tolist()
    theta = 2 * math.pi * u[0]
    phi = math.radians(15 + 60 * u[1])  # elevation on the upper hemisphere
    eye = radius * np.array([math.cos(theta) * math.cos(phi), math.sin(theta) * math.cos(phi), math.sin(phi)])
    return look_at_camera(eye, width, height, NERF_SYNTHETIC_FOVX)


def bicycle_camera(seed: int = 0, width: int = 1245, height: int = 825, radius: float = 5.0) -> SynthCamera:
    g = torch.Generator().manual_seed(2000 + seed)
    u = torch.rand(2, generator=g).tolist()
    theta = 2 * math.pi * u[0]
    phi = math.radians(10 + 25 * u[1])
    eye = radius * np.array([math.cos(theta) * math.cos(phi), math.sin(theta) * math.cos(phi), math.sin(phi)])
    return look_at_camera(eye, width, height, fovx=0.9)


def ring_cameras(n: int, width: int, height: int, radius: float, fovx: float, seed: int = 0):
    g = torch.Generator().manual_seed(3000 + seed)
    elev = (torch.rand(n, generator=g) * 30 + 10).tolist()
    cams = []
    for i in range(n):
        th, ph = 2 * math.pi * i / n, math.radians(elev[i])
        eye = radius * np.array([math.cos(th) * math.cos(ph), math.sin(th) * math.cos(ph), math.sin(ph)])
        cams.append(look_at_camera(eye, width, height, fovx))
    return cams


def random_cloud(P: int, seed: int = 0, extent: float = 1.3, log_scale_mean: float = math.log(0.01),
                 sh_coeffs: int = 16) -> dict:
    """C1 law of SURVEY.md 8d: free Gaussians, `shs` + `scales` + `rotations` mode."""
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand(P, 3, generator=g) * 2 - 1) * extent
    ls = log_scale_mean + 0.6 * torch.randn(P, 1, generator=g) + 0.3 * torch.randn(P, 3, generator=g)
    scales = torch.exp(ls)
    q = torch.randn(P, 4, generator=g)
    rotations = q / q.norm(dim=1, keepdim=True)
    opacities = torch.sigmoid(2.0 * torch.randn(P, 1, generator=g))
    shs = 0.3 * torch.randn(P, sh_coeffs, 3, generator=g)
    shs[:, 0, :] = torch.randn(P, 3, generator=g)
    return {"means3D": xyz.contiguous(), "scales": scales.contiguous(), "rotations": rotations.contiguous(),
            "opacities": opacities.contiguous(), "shs": shs.contiguous()}


def icosphere(subdiv: int, radius: float = 1.0):
    """Closed triangle mesh: (verts [V,3] float32, faces [F,3] int64); F = 20 * 4**subdiv."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
         (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
         (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = np.array(v, dtype=np.float64)
    verts /= np.linalg.norm(verts, axis=1, keepdims=True)
    faces = np.array(f, dtype=np.int64)
    for _ in range(subdiv):
        e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
        e.sort(axis=1)
        ue, inv = np.unique(e, axis=0, return_inverse=True)
        mid = verts[ue[:, 0]] + verts[ue[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        nF, base = faces.shape[0], verts.shape[0]
        m01, m12, m20 = base + inv[:nF], base + inv[nF:2 * nF], base + inv[2 * nF:]
        verts = np.concatenate([verts, mid], axis=0)
        a, b, c = faces[:, 0], faces[:, 1], faces[:, 2]
        faces = np.concatenate([np.stack([a, m01, m20], 1), np.stack([b, m12, m01], 1),
                                np.stack([c, m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    return (torch.tensor(verts * radius, dtype=torch.float32).contiguous(),
            torch.tensor(faces, dtype=torch.int64).contiguous())


def jittered_sphere_mesh(n_faces_target: int, seed: int = 0, radius: float = 1.0, jitter: float = 0.02):
    """C2 mesh: icosphere subdivided to >= n_faces_target faces, vertices jittered by 2 % of edge."""
    sub = 0
    while 20 * 4 ** sub < n_faces_target:
        sub += 1
    verts, faces = icosphere(sub, radius)
    g = torch.Generator().manual_seed(4000 + seed)
    edge = (verts[faces[:, 1]] - verts[faces[:, 0]]).norm(dim=1).mean()
    verts = verts + jitter * edge * torch.randn(verts.shape, generator=g)
    return verts.contiguous(), faces


def barycentric_layout(k: int):
    """geo/mesh_utils.py:16-40 (generate_barycentric_v2): (bc[k,3] float32, rad_base)."""
    s3 = 3 ** 0.5
    if k == 1:
        bc, r = [[1 / 3, 1 / 3, 1 / 3]], 1.0 / (2.0 * s3)
    elif k == 3:
        a, b = (3 - s3) / 6, s3 / 3
        bc, r = [[a, a, b], [a, b, a], [b, a, a]], 1.0 / (2.0 + 2.0 * s3)
    elif k == 6:
        bc = [[2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3],
              [1 / 6, 5 / 12, 5 / 12], [5 / 12, 1 / 6, 5 / 12], [5 / 12, 5 / 12, 1 / 6]]
        r = 1 / (4.0 + 2.0 * s3)
    else:
        raise NotImplementedError
    return torch.tensor(bc, dtype=torch.float32), float(r)


def mesh_bound_inputs(n_faces_target: int = 50_000, k: int = 6, seed: int = 1, sh_coeffs: int = 16) -> dict:
    """C2 law: mesh + per-Gaussian SH features [P,3,M] (DMGS `features` layout), opacity 0.9999."""
    verts, faces = jittered_sphere_mesh(n_faces_target, seed)
    P = faces.shape[0] * k
    g = torch.Generator().manual_seed(5000 + seed)
    features = 0.3 * torch.randn(P, 3, sh_coeffs, generator=g)
    bc, rad_base = barycentric_layout(k)
    return {"verts": verts, "faces": faces, "bc": bc, "rad_base": rad_base, "k": k,
            "features": features.contiguous(), "opacities": torch.full((P, 1), 0.9999),
            "spatial_lr_scale": 4.43, "scale_factor": math.atanh(0.5)}
