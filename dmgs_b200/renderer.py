"""`render` / `render_dyn` with the reference's signatures and return dict, on the fused path.

Mirrors gaussian_renderer/__init__.py:18-101 (`render`, stage 1 / stage 3: a model object with
`get_xyz`, `get_opacity`, `get_scaling`, `get_rotation`, `get_features`, `get_covariance`) and
:104-193 (`render_dyn`, stage 2: the `gs_info` dict of scene/gaussian_geo_model_mlp_flex.py:321-334).
The difference is where the SH colour is evaluated: when `pipe.convert_SHs_python` is set the
reference runs `sigmoid(eval_sh(...))` in PyTorch (:74-78, :166-170) and hands `colors_precomp` to
the rasteriser; here the same basis + activation run inside the preprocess kernel
(`sh_activation="sigmoid"`), so no [P,3] colour tensor and no [P,3,16] transpose copy is materialised
and `dL/dfeatures` comes straight out of the fused backward.  Return dict keys are unchanged:
"render", "viewspace_points", "visibility_filter", "radii".
"""
from __future__ import annotations

import math

import torch

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def _settings(cam, bg_color, scaling_modifier, sh_degree, debug):
    dev = bg_color.device
    return GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width),
        tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=cam.world_view_transform.to(dev),
        projmatrix=cam.full_proj_transform.to(dev), sh_degree=int(sh_degree),
        campos=cam.camera_center.to(dev), prefiltered=False, debug=bool(debug))


def _package(image, radii, screenspace_points):
    return {"render": image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii}


def _screenspace_like(xyz):
    pts = torch.zeros_like(xyz, requires_grad=True)
    try:
        pts.retain_grad()
    except RuntimeError:
        pass
    return pts


def render_dyn(viewpoint_camera, gs_info: dict, pipe, bg_color: torch.Tensor, scaling_modifier=1.0,
               override_color=None):
    """Stage-2 render (gaussian_renderer/__init__.py:104-193).  `gs_info["features"]` is [P,3,16]."""
    xyz, opacity = gs_info["xyz"], gs_info["opacity"]
    if not pipe.compute_cov3D_python:
        raise NotImplementedError  # as the reference (:158-159): stage 2 always feeds cov3D_precomp
    means2D = _screenspace_like(xyz)
    rs = _settings(viewpoint_camera, bg_color, scaling_modifier, gs_info["active_sh_degree"], pipe.debug)
    kw = dict(means3D=xyz, means2D=means2D, opacities=opacity, scales=None, rotations=None,
              cov3D_precomp=gs_info["covariance"])
    if override_color is not None:
        ras = GaussianRasterizer(raster_settings=rs)
        image, radii = ras(shs=None, colors_precomp=override_color, **kw)
    else:
        M = (gs_info["max_sh_degree"] + 1) ** 2
        feats = gs_info["features"].view(-1, 3, M)
        act = "sigmoid" if pipe.convert_SHs_python else "clamp"
        ras = GaussianRasterizer(raster_settings=rs, sh_activation=act, sh_layout="P3M")
        image, radii = ras(shs=feats, colors_precomp=None, **kw)
    return _package(image, radii, means2D)


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None):
    """Stage-1 / stage-3 render (gaussian_renderer/__init__.py:18-101).  `pc.get_features` is [P,16,3]."""
    xyz = pc.get_xyz
    means2D = _screenspace_like(xyz)
    rs = _settings(viewpoint_camera, bg_color, scaling_modifier, pc.active_sh_degree, pipe.debug)
    kw = dict(means3D=xyz, means2D=means2D, opacities=pc.get_opacity)
    if pipe.compute_cov3D_python:
        kw.update(scales=None, rotations=None, cov3D_precomp=pc.get_covariance(scaling_modifier))
    else:
        kw.update(scales=pc.get_scaling, rotations=pc.get_rotation, cov3D_precomp=None)
    if override_color is not None:
        ras = GaussianRasterizer(raster_settings=rs)
        image, radii = ras(shs=None, colors_precomp=override_color, **kw)
    else:
        # the reference transposes get_features to [P,3,M] for its python path (:75); the kernel
        # reads the [P,M,3] layout directly, so no copy is made in either case
        act = "sigmoid" if pipe.convert_SHs_python else "clamp"
        ras = GaussianRasterizer(raster_settings=rs, sh_activation=act, sh_layout="PM3")
        image, radii = ras(shs=pc.get_features, colors_precomp=None, **kw)
    return _package(image, radii, means2D)
