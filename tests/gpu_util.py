"""GPU-side helpers for the parity tests: run the CUDA path through the C ABI and return numpy."""
import math

import numpy as np
import torch

from dmgs_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_forward, rasterize_backward


def settings_for(cam, bg, sh_degree=3, scale_modifier=1.0, debug=False, device="cuda"):
    return GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width), tanfovx=math.tan(cam.FoVx * 0.5),
        tanfovy=math.tan(cam.FoVy * 0.5), bg=torch.tensor(bg, dtype=torch.float32, device=device),
        scale_modifier=scale_modifier, viewmatrix=cam.world_view_transform.to(device),
        projmatrix=cam.full_proj_transform.to(device), sh_degree=sh_degree, campos=cam.camera_center.to(device),
        prefiltered=False, debug=debug)


def gpu_forward(settings, means3D, opacities, sh_layout=0, sh_activation=0, **kw):
    d = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.float32).contiguous().cuda()
    color, radii, st = rasterize_forward(settings, d(means3D), d(opacities), d(kw.get("shs")), d(kw.get("colors_precomp")),
                                         d(kw.get("scales")), d(kw.get("rotations")), d(kw.get("cov3D_precomp")),
                                         sh_layout, sh_activation)
    torch.cuda.synchronize()
    return color, radii, st


def state_numpy(st):
    g = {k: v.cpu().numpy() for k, v in st.geom_arrays().items()}
    b = {k: v.cpu().numpy() for k, v in st.binning_arrays().items()}
    im = {k: v.cpu().numpy() for k, v in st.image_arrays().items()}
    keys = st.sorted_keys().cpu().numpy().view(np.uint64)
    return g, b, im, keys
