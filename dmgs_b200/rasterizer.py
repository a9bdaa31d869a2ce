"""`GaussianRasterizationSettings` / `GaussianRasterizer` -- same Python surface as the
`diff_gaussian_rasterization` package DMGS imports (gaussian_renderer/__init__.py:14, settings
tuple :36-49, call :86-94), backed by libdmgs_raster.so through ctypes.

Extensions that keep the drop-in intact (all optional, default = upstream behaviour):
  * ``GaussianRasterizer(raster_settings, sh_activation="clamp"|"sigmoid", sh_layout="PM3"|"P3M")``
    lets DMGS's python ``eval_sh`` + ``sigmoid`` (gaussian_renderer/__init__.py:74-78, :166-170) run
    inside preprocess: pass ``shs=features`` instead of ``colors_precomp``.
  * ``rasterizer.last`` keeps the state buffers of the last forward for inspection (tests).
  * ``configure(async_binning=True)`` removes the one host synchronisation of the forward (the
    read-back of the instance count that sizes the binning buffer, which upstream also performs):
    the buffer is sized from the counts seen so far for the same (P, W, H) plus slack, the count
    stays on the device, and ``check_async()`` -- called at the caller's next natural sync point,
    e.g. right after ``loss.item()`` -- reports whether any frame since the last check overflowed
    its buffer (such a frame rendered as background with zero gradients and must be repeated).
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib as L


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


_HOST_CACHE: dict = {}

# ---- optional sync-free binning (see the module docstring) ------------------------------------------
_ASYNC = {"on": False, "slack": 1.25, "capacity": {}, "pending": []}


def configure(async_binning: Optional[bool] = None, capacity_slack: Optional[float] = None):
    """async_binning: size the binning buffer from earlier frames instead of reading the instance
    count back (default False = upstream behaviour).  capacity_slack: head-room factor (default 1.25)."""
    if async_binning is not None:
        _ASYNC["on"] = bool(async_binning)
    if capacity_slack is not None:
        _ASYNC["slack"] = float(capacity_slack)
    if not _ASYNC["on"]:
        _ASYNC["pending"].clear()


def check_async() -> bool:
    """True if every async-binned frame since the last call fitted its buffer.  Synchronises (one
    small device->host copy).  On overflow the capacity for that shape is raised and False is
    returned: the caller repeats the step."""
    pend, _ASYNC["pending"] = _ASYNC["pending"], []
    if not pend:
        return True
    flags = torch.cat([f for _, f, _ in pend]).cpu().tolist()
    counts = torch.cat([n for _, _, n in pend]).cpu().tolist()
    ok = True
    for (key, _, _), need, count in zip(pend, flags, counts):
        cap = _ASYNC["capacity"].get(key, 0)
        _ASYNC["capacity"][key] = max(cap, int(max(need, count) * _ASYNC["slack"]) + 4096)
        ok = ok and need == 0
    return ok


def _host_values(tensors):
    """Flat float lists of the small camera tensors (bg, viewmatrix, projmatrix, campos).

    Values are cached per tensor OBJECT (weak reference + in-place version counter), never per
    data pointer: the caching allocator hands the same address to unrelated tensors.  All misses
    of one call share a single device->host copy."""
    out, miss = [None] * len(tensors), []
    for i, t in enumerate(tensors):
        hit = _HOST_CACHE.get(id(t))
        if hit is not None and hit[0]() is t and hit[1] == t._version:
            out[i] = hit[2]
        else:
            miss.append(i)
    if miss:
        flat = torch.cat([tensors[i].detach().reshape(-1).float() for i in miss]).cpu().tolist()
        o = 0
        for i in miss:
            t = tensors[i]
            n = t.numel()
            vals = flat[o:o + n]
            o += n
            out[i] = vals
            key = id(t)
            try:
                _HOST_CACHE[key] = (weakref.ref(t, lambda _r, k=key: _HOST_CACHE.pop(k, None)), t._version, vals)
            except TypeError:
                pass
    return out


def make_params(settings: GaussianRasterizationSettings, P: int, sh_coeffs: int, sh_layout: int,
                sh_activation: int) -> L.DmgsParams:
    prm = L.DmgsParams()
    prm.P = int(P)
    prm.sh_degree = int(settings.sh_degree)
    prm.sh_coeffs = int(sh_coeffs)
    prm.image_width, prm.image_height = int(settings.image_width), int(settings.image_height)
    prm.sh_layout, prm.sh_activation = int(sh_layout), int(sh_activation)
    prm.debug = int(bool(settings.debug))
    prm.tanfovx, prm.tanfovy = float(settings.tanfovx), float(settings.tanfovy)
    prm.scale_modifier = float(settings.scale_modifier)
    bg, view, proj, campos = _host_values([settings.bg, settings.viewmatrix, settings.projmatrix, settings.campos])
    prm.bg[:], prm.viewmatrix[:], prm.projmatrix[:], prm.campos[:] = bg, view, proj, campos
    return prm


def _f32c(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class RasterState:
    """Caller-owned state of one forward (the geom / binning / image byte buffers)."""

    def __init__(self, prm, geom, binning, image, num_rendered, radii, count_dev=None):
        self.prm, self.geom, self.binning, self.image = prm, geom, binning, image
        # layout_R sizes the binning layout in every later C call: the instance count on the synchronous
        # path, the buffer capacity on the async path (where the real count stays on the device)
        self.layout_R, self.radii, self._count_dev = num_rendered, radii, count_dev
        self._count = None if count_dev is not None else num_rendered

    @property
    def num_rendered(self) -> int:
        if self._count is None:
            self._count = int(self._count_dev.item())
        return self._count

    # typed views for the parity tests -------------------------------------------------
    def _view(self, buf, off, dtype, shape):
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return buf[off:off + nbytes].view(dtype).view(*shape)

    def geom_arrays(self):
        P = self.prm.P
        o = (C.c_int64 * 9)()
        L.lib().dmgs_geom_layout(P, o)
        v = self._view
        return {"depths": v(self.geom, o[0], torch.float32, (P,)), "rec": v(self.geom, o[1], torch.float32, (P, 8)),
                "rgb": v(self.geom, o[2], torch.float32, (P, 4)), "clamped": v(self.geom, o[3], torch.uint8, (P,)),
                "cov3D": v(self.geom, o[4], torch.float32, (P, 6)), "tiles_touched": v(self.geom, o[5], torch.int32, (P,)),
                "rect": v(self.geom, o[6], torch.int16, (P, 4)), "order": v(self.geom, o[7], torch.int32, (P,)),
                "offsets": v(self.geom, o[8], torch.int32, (P,))}

    def binning_arrays(self):
        R, W, H = self.num_rendered, self.prm.image_width, self.prm.image_height
        T = ((W + 15) // 16) * ((H + 15) // 16)
        o = (C.c_int64 * 3)()
        L.lib().dmgs_binning_layout(self.prm.P, self.layout_R, W, H, o)
        v = self._view
        return {"tiles": v(self.binning, o[0], torch.int32, (R,)), "gidx": v(self.binning, o[1], torch.int32, (R,)),
                "ranges": v(self.binning, o[2], torch.int32, (T, 2))}

    def image_arrays(self):
        W, H = self.prm.image_width, self.prm.image_height
        o = (C.c_int64 * 2)()
        L.lib().dmgs_image_layout(W, H, o)
        v = self._view
        return {"final_T": v(self.image, o[0], torch.float32, (H, W)),
                "n_contrib": v(self.image, o[1], torch.int32, (H, W))}

    def sorted_keys(self):
        R = self.num_rendered
        keys = torch.empty(max(R, 1), dtype=torch.int64, device=self.geom.device)
        if self.layout_R != R:
            raise RuntimeError("sorted_keys() is an inspection helper of the synchronous path")
        L.check(L.lib().dmgs_sorted_keys(L.ptr(self.geom), L.ptr(self.binning), self.prm.P, R, self.prm.image_width,
                                         self.prm.image_height, L.ptr(keys), _stream()), "dmgs_sorted_keys")
        return keys[:R]


def rasterize_forward(settings, means3D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp,
                      sh_layout=0, sh_activation=0, stage_hook=None):
    """Runs the three forward stages; returns (color [3,H,W], radii [P] int32, RasterState)."""
    dev = means3D.device
    if dev.type != "cuda":
        raise RuntimeError("dmgs_b200 rasteriser needs CUDA tensors; there is no CPU path")
    lib = L.lib()
    P = int(means3D.shape[0])
    H, W = int(settings.image_height), int(settings.image_width)
    M = 0
    if shs is not None:
        M = int(shs.shape[1] if sh_layout == 0 else shs.shape[2])
    prm = make_params(settings, P, M, sh_layout, sh_activation)
    stream = _stream()
    radii = torch.zeros(P, dtype=torch.int32, device=dev)
    color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
    geom = torch.empty(lib.dmgs_geom_bytes(P), dtype=torch.uint8, device=dev)
    image = torch.empty(lib.dmgs_image_bytes(W, H), dtype=torch.uint8, device=dev)
    nr = torch.zeros(1, dtype=torch.int32, device=dev)
    L.check(lib.dmgs_preprocess_forward(C.byref(prm), L.ptr(means3D), L.ptr(scales), L.ptr(rotations),
                                        L.ptr(cov3D_precomp), L.ptr(opacities), L.ptr(shs), L.ptr(colors_precomp),
                                        L.ptr(radii), L.ptr(geom), L.ptr(nr), stream), "dmgs_preprocess_forward")
    if stage_hook is not None:
        stage_hook("preprocess_sort_scan")
    key = (dev.index, P, W, H)
    cap = _ASYNC["capacity"].get(key) if _ASYNC["on"] else None
    count_dev = None
    if cap is not None:
        # sync-free: buffer sized from earlier frames of this shape; the count stays on the device
        R = int(cap)
        binning = torch.empty(lib.dmgs_binning_bytes(P, R, W, H), dtype=torch.uint8, device=dev)
        flag = torch.empty(1, dtype=torch.int32, device=dev)
        rc = lib.dmgs_bin_forward_async(C.byref(prm), L.ptr(geom), R, L.ptr(binning), L.ptr(flag), stream)
        if rc == -9:  # image too large for the placement path: synchronous from now on
            _ASYNC["capacity"].pop(key, None)
            cap = None
        else:
            L.check(rc, "dmgs_bin_forward_async")
            _ASYNC["pending"].append((key, flag, nr))
            count_dev = nr
    if cap is None:
        R = int(nr.item())  # the one host read-back of the forward (upstream does the same after its scan)
        if _ASYNC["on"]:
            _ASYNC["capacity"][key] = max(_ASYNC["capacity"].get(key, 0), int(R * _ASYNC["slack"]) + 4096)
        binning = torch.empty(lib.dmgs_binning_bytes(P, R, W, H), dtype=torch.uint8, device=dev)
        L.check(lib.dmgs_bin_forward(C.byref(prm), L.ptr(geom), R, L.ptr(binning), stream), "dmgs_bin_forward")
    if stage_hook is not None:
        stage_hook("binning")
    L.check(lib.dmgs_blend_forward(C.byref(prm), L.ptr(geom), L.ptr(binning), R, L.ptr(color), L.ptr(image), stream),
            "dmgs_blend_forward")
    if stage_hook is not None:
        stage_hook("blend_fwd")
    return color, radii, RasterState(prm, geom, binning, image, R, radii, count_dev)


def rasterize_backward(state: RasterState, grad_color, means3D, shs, scales, rotations, cov3D_precomp,
                       want_colors_precomp, stage_hook=None, accumulate_into=None, sh_record=None):
    """accumulate_into: optional dict of preallocated gradient tensors (keys means3D, means2D, opacities,
    colors_precomp, shs, scales, rotations, cov3D_precomp) that the gradients are ADDED to.
    sh_record: with accumulate_into, a float32 [P,4] tensor that receives this view's deferred SH gradient
    record {g.r, g.g, g.b, seen} instead of the [P,16,3] rows being read-modify-written
    (dmgs_preprocess_backward accumulate = 2; `sh_grad_expand` forms the rows once per step)."""
    lib = L.lib()
    prm, P, dev = state.prm, state.prm.P, means3D.device
    z = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    g_means3D, g_means2D, g_op = z(P, 3), z(P, 3), z(P, 1)
    g_col = z(P, 3) if want_colors_precomp else None
    g_shs = torch.empty_like(shs) if shs is not None else None
    g_scales = z(P, 3) if scales is not None else None
    g_rots = z(P, 4) if rotations is not None else None
    g_cov = z(P, 6) if cov3D_precomp is not None else None
    if accumulate_into is not None:
        a = accumulate_into
        g_means3D, g_means2D, g_op = a["means3D"], a["means2D"], a["opacities"]
        g_col = a.get("colors_precomp") if want_colors_precomp else None
        g_shs = a.get("shs") if shs is not None else None
        if sh_record is not None and shs is not None:
            if sh_record.shape != (P, 4) or sh_record.dtype != torch.float32 or not sh_record.is_contiguous():
                raise ValueError("sh_record must be a contiguous float32 [P,4] tensor")
            g_shs = sh_record
        g_scales = a.get("scales") if scales is not None else None
        g_rots = a.get("rotations") if rotations is not None else None
        g_cov = a.get("cov3D_precomp") if cov3D_precomp is not None else None
    scratch = torch.empty(lib.dmgs_backward_scratch_bytes(P), dtype=torch.uint8, device=dev)
    stream = _stream()
    L.check(lib.dmgs_blend_backward(C.byref(prm), L.ptr(state.geom), L.ptr(state.binning), L.ptr(state.image),
                                    state.layout_R, L.ptr(grad_color), L.ptr(scratch), stream),
            "dmgs_blend_backward")
    if stage_hook is not None:
        stage_hook("blend_bwd")
    L.check(lib.dmgs_preprocess_backward(C.byref(prm), L.ptr(means3D), L.ptr(scales), L.ptr(rotations),
                                         L.ptr(cov3D_precomp), L.ptr(shs), L.ptr(state.radii), L.ptr(state.geom),
                                         L.ptr(scratch), L.ptr(g_means3D), L.ptr(g_means2D), L.ptr(g_op), L.ptr(g_col),
                                         L.ptr(g_shs), L.ptr(g_scales), L.ptr(g_rots), L.ptr(g_cov),
                                         (2 if (sh_record is not None and shs is not None) else 1)
                                         if accumulate_into is not None else 0, stream),
            "dmgs_preprocess_backward")
    if stage_hook is not None:
        stage_hook("preprocess_bwd")
    return g_means3D, g_means2D, g_shs, g_col, g_op, g_scales, g_rots, g_cov


def sh_grad_expand(records, camera_centers, means3D, shs, sh_degree, out, out_means3D, sh_layout=0, accumulate=False):
    """Forms dL/dsh and the view-direction term of dL/dmeans3D from the deferred records of a step's views
    (libdmgs_raster.so: dmgs_sh_grad_expand).  records: float32 [V,P,4] (views along dim 0); camera_centers: V
    tensors/sequences of 3 floats (the campos of each view's settings, in the same order); shs: the coefficients;
    out: the [P,M,3] (sh_layout 0) or [P,3,M] (1) gradient tensor, overwritten (accumulate=False) or added to;
    out_means3D: [P,3], always added to."""
    V, P = int(records.shape[0]), int(records.shape[1])
    M = int(out.shape[1] if sh_layout == 0 else out.shape[2])
    cams = []
    for c in camera_centers:
        cams.extend(_host_values([c])[0] if isinstance(c, torch.Tensor) else [float(x) for x in c])
    if len(cams) != 3 * V:
        raise ValueError("one camera centre per view")
    arr = (C.c_float * (3 * V))(*cams)
    stride = int(records.stride(0)) if V > 1 else P * 4
    L.check(L.lib().dmgs_sh_grad_expand(P, int(sh_degree), M, int(sh_layout), V, arr, L.ptr(means3D), L.ptr(shs),
                                        L.ptr(records), stride, L.ptr(out), L.ptr(out_means3D), int(bool(accumulate)),
                                        _stream()), "dmgs_sh_grad_expand")
    return out


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, sh_layout, sh_activation, holder):
        means3D_c, op_c = _f32c(means3D), _f32c(opacities)
        sh_c, col_c = _f32c(sh), _f32c(colors_precomp)
        sc_c, rot_c, cov_c = _f32c(scales), _f32c(rotations), _f32c(cov3Ds_precomp)
        if means3D_c is None:  # P == 0
            H, W = int(raster_settings.image_height), int(raster_settings.image_width)
            color = raster_settings.bg.float().view(3, 1, 1).expand(3, H, W).contiguous()
            ctx.state = None
            return color, torch.zeros(0, dtype=torch.int32, device=color.device)
        color, radii, state = rasterize_forward(raster_settings, means3D_c, op_c, sh_c, col_c, sc_c, rot_c, cov_c,
                                                sh_layout, sh_activation)
        ctx.state = state
        ctx.save_for_backward(means3D_c, sh_c, sc_c, rot_c, cov_c)
        ctx.has_col = col_c is not None
        if holder is not None:
            holder.last = state
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, _grad_radii):
        if ctx.state is None:
            return (None,) * 12
        means3D, sh, scales, rotations, cov = ctx.saved_tensors
        g = rasterize_backward(ctx.state, grad_color.contiguous().float(), means3D, sh, scales, rotations, cov,
                               ctx.has_col)
        g_means3D, g_means2D, g_shs, g_col, g_op, g_scales, g_rots, g_cov = g
        return (g_means3D, g_means2D, g_shs, g_col, g_op, g_scales, g_rots, g_cov, None, None, None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, sh_layout=0, sh_activation=0, holder=None):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, sh_layout, sh_activation, holder)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings, sh_activation: str = "clamp", sh_layout: str = "PM3"):
        super().__init__()
        self.raster_settings = raster_settings
        self.sh_activation = {"clamp": 0, "sigmoid": 1}[sh_activation]
        self.sh_layout = {"PM3": 0, "P3M": 1}[sh_layout]
        self.last: Optional[RasterState] = None

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            pos = _f32c(positions)
            P = int(positions.shape[0])
            vis = torch.zeros(P, dtype=torch.uint8, device=positions.device)
            if P:
                view = rs.viewmatrix.float().contiguous()
                proj = rs.projmatrix.float().contiguous()
                L.check(L.lib().dmgs_mark_visible(P, L.ptr(pos), L.ptr(view), L.ptr(proj), L.ptr(vis), _stream()),
                        "dmgs_mark_visible")
            return vis.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   self.raster_settings, self.sh_layout, self.sh_activation, self)
