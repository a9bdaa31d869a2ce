"""View-frustum test and visible-face compaction on the device (libdmgs_raster.so: dmgs_in_frustum).

Host-side mirror of the reference's ``in_frustum`` helpers and of how stage 3 uses them:
  * ``in_frustum(proj_matrix, verts)``                                   <- scene/gaussian_geo_model_finetune.py:33-48
  * ``in_frustum(proj_matrix, verts, cube_len, piece_id, n_piece)``      <- scene/gaussian_geo_model_mlp_flex_colmap.py:32-76
  * ``cull_faces(proj_matrix, verts, faces, gs_per_face)`` -> (face_mask, gs_mask, faces[face_mask])
                                                                          <- finetune.py:405-409
``cull_faces`` never materialises the centroids: one kernel gathers the three vertices, tests the
centroid and counts, a block scan and one scatter produce ``faces[mask]`` in order.  Like boolean
indexing in PyTorch it reads the visible count back to the host (the output shape depends on it).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from .rasterizer import _host_values


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _proj16(proj_matrix: torch.Tensor):
    (vals,) = _host_values([proj_matrix])
    if len(vals) != 16:
        raise ValueError("proj_matrix must be [4,4]")
    return (C.c_float * 16)(*vals)


def in_frustum(proj_matrix: torch.Tensor, verts: torch.Tensor, cube_len: Optional[float] = None, piece_id: int = -1,
               n_piece: int = 1) -> torch.Tensor:
    """bool [N]: the reference's mask, bit for bit (fp32; see csrc/frustum.cu for the operation order)."""
    if verts.device.type != "cuda":
        raise RuntimeError("dmgs_b200 frustum test needs CUDA tensors; there is no CPU path")
    if piece_id >= 0 and not ((n_piece == 2 and piece_id < 2) or (n_piece == 4 and piece_id < 4)):
        raise NotImplementedError  # as the reference (colmap.py:57, :71, :73)
    pts = verts.detach().float().contiguous()
    N = int(pts.shape[0])
    mask = torch.empty(N, dtype=torch.bool, device=pts.device)
    L.check(L.lib().dmgs_in_frustum(N, _proj16(proj_matrix), float(cube_len or 0.0), int(cube_len is not None),
                                    int(piece_id), int(n_piece), L.ptr(pts), None, L.ptr(mask), None, None, None, None,
                                    _stream()), "dmgs_in_frustum")
    return mask


def cull_faces(proj_matrix: torch.Tensor, verts: torch.Tensor, faces: torch.Tensor, gs_per_face: int = 1):
    """(face_mask [F] bool, gs_mask [F*k] bool, faces[face_mask] [F',3] int64) for the camera's frustum,
    testing the face centroids ``verts[faces].mean(dim=1)`` (finetune.py:405-409)."""
    if verts.device.type != "cuda":
        raise RuntimeError("dmgs_b200 frustum test needs CUDA tensors; there is no CPU path")
    v = verts.detach().float().contiguous()
    f = faces.to(torch.int64).contiguous()
    F = int(f.shape[0])
    lib = L.lib()
    mask = torch.empty(F, dtype=torch.bool, device=v.device)
    out = torch.empty(F, 3, dtype=torch.int64, device=v.device)
    count = torch.zeros(1, dtype=torch.int32, device=v.device)
    scratch = torch.empty(lib.dmgs_frustum_scratch_bytes(F), dtype=torch.uint8, device=v.device)
    L.check(lib.dmgs_in_frustum(F, _proj16(proj_matrix), 0.0, 0, -1, 1, L.ptr(v), L.ptr(f), L.ptr(mask), L.ptr(out), None,
                                L.ptr(count), L.ptr(scratch), _stream()), "dmgs_in_frustum")
    n = int(count.item())
    gs_mask = torch.repeat_interleave(mask, int(gs_per_face)) if gs_per_face != 1 else mask
    return mask, gs_mask, out[:n]
