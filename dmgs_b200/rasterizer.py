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
from collections import OrderedDict
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


class BinningOverflowError(RuntimeError):
    """A frame rendered with sync-free binning did not fit its binning buffer (it rendered as background with
    zero gradients).  The capacity for its shape has been raised; repeat the step."""


# ---- optional sync-free binning (see the module docstring) ------------------------------------------
# capacity: instances the binning buffer is sized for, per (device, P, W, H); no_async: shapes that cannot use
# the placement path (> 16384 tiles) and stay synchronous; pending: the probes (count + overflow flag on their
# way to pinned memory) of the frames nobody has looked at yet, per device.
_ASYNC = {"on": False, "slack": 1.25, "capacity": {}, "no_async": set(), "pending": {}}
_MAX_PENDING = 256


# ---- optional fused gradient accumulation of the autograd module (gradient-accumulation loops) -----------------------
_FUSED_ACC = {"on": False}


def configure(async_binning: Optional[bool] = None, capacity_slack: Optional[float] = None,
              accumulate_grad_in_place: Optional[bool] = None):
    """async_binning: size the binning buffer from earlier frames instead of reading the instance
    count back (default False = upstream behaviour).  capacity_slack: head-room factor (default 1.25).

    Sync-free frames are verified, never trusted silently: every frame copies its instance count and
    overflow flag to pinned host memory behind an event; `rasterize_backward` (and the autograd
    backward) waits for that event -- long complete by then -- and raises `BinningOverflowError` if
    the frame overflowed, `check_async()` polls all frames since the last call.

    accumulate_grad_in_place (default False): when the autograd module's backward finds that an input is a leaf whose
    `.grad` already holds a gradient (the second and later views of a gradient-accumulation step), the kernels ADD
    into that `.grad` and autograd receives None for the input -- instead of writing a fresh [P, ...] gradient that
    autograd's AccumulateGrad then adds with one more pass over both tensors (H0: 98 us of a 1089 us frame).  Only
    meaningful under `loss.backward()`: `torch.autograd.grad`, tensor hooks and `create_graph` never see those
    gradients, so it is opt-in (the same trade as Megatron's gradient-accumulation fusion)."""
    if async_binning is not None:
        _ASYNC["on"] = bool(async_binning)
    if capacity_slack is not None:
        _ASYNC["slack"] = float(capacity_slack)
    if accumulate_grad_in_place is not None:
        _FUSED_ACC["on"] = bool(accumulate_grad_in_place)
    if not _ASYNC["on"]:
        _ASYNC["pending"].clear()


def _raise_capacity(key, need):
    cap = _ASYNC["capacity"].get(key, 0)
    _ASYNC["capacity"][key] = max(cap, int(need * _ASYNC["slack"]) + 4096)


def check_async(device=None) -> bool:
    """True if every sync-free frame since the last call fitted its buffer (all devices, or `device`).
    Waits for those frames' events (one per frame, normally complete already).  On overflow the capacity
    for that shape is raised and False is returned: the caller repeats the step."""
    ok = True
    for dev_index in list(_ASYNC["pending"]):
        if device is not None and torch.device(device).index not in (None, dev_index):
            continue
        pend, _ASYNC["pending"][dev_index] = _ASYNC["pending"][dev_index], []
        for probe in pend:
            if not probe.resolve():
                ok = False
    return ok


class _Probe:
    """(instance count, overflow flag) of one sync-free frame on its way to pinned host memory."""

    __slots__ = ("key", "capacity", "host", "event", "count", "need")

    def __init__(self, key, capacity, meta, stream):
        self.key, self.capacity, self.count, self.need = key, capacity, None, None
        self.host = torch.empty(2, dtype=torch.int32, pin_memory=True)
        self.host.copy_(meta, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record(stream)

    def resolve(self) -> bool:
        """Waits for the copy (normally long complete); raises the shape's capacity; True if the frame fitted."""
        if self.count is None:
            self.event.synchronize()
            self.count, self.need = int(self.host[0]), int(self.host[1])
            self.host = self.event = None
            _raise_capacity(self.key, max(self.count, self.need))
        return self.need == 0


# ---- frames captured into CUDA graphs ---------------------------------------------------------------------
# A sync-free frame has a fixed launch sequence (every data-dependent size stays on the device), so a whole view --
# forward, loss gradient, backward -- can be captured once and replayed with one graph launch instead of ~25 kernel
# launches and their host glue (multiview.ViewStreams.capture).  The eager probe above cannot live in a graph (it
# allocates pinned memory and records an event the host later waits for), so a captured frame reports through a
# CaptureProbe: every replay bumps a sequence number on the device and copies {count, needed, sequence} to pinned
# memory right after the binning; the host polls that memory -- no CUDA synchronisation at all.
_CAPTURE = {"probe": None}


class CaptureProbe:
    """{instance count, overflow: count needed, replay number} of a frame that lives in a CUDA graph."""

    def __init__(self, device):
        self.dev = torch.zeros(4, dtype=torch.int32, device=device)
        # page-locked through cudaHostRegister, NOT torch's caching host allocator: that allocator records an event
        # per non-blocking copy to track the block, and an event recorded under capture can never be queried
        self.host = torch.zeros(1024, dtype=torch.int32)[:4]
        self._registered = False
        try:
            rc = torch.cuda.cudart().cudaHostRegister(self.host.data_ptr(), 16, 0)
            self._registered = int(rc) == 0
        except Exception:
            self._registered = False
        if not self._registered:
            self.host = torch.zeros(4, dtype=torch.int32).pin_memory()
        self._np = self.host.numpy()  # shares the pinned storage: polling is a plain memory read
        self.key, self.capacity, self.replays, self.seen = None, None, 0, 0
        self.count, self.need = None, None
        self.keep = []  # the frame's workspace: its addresses are baked into the graph, it never returns to the pool

    def __del__(self):
        if getattr(self, "_registered", False):
            try:
                torch.cuda.cudart().cudaHostUnregister(self.host.data_ptr())
            except Exception:
                pass

    def record(self, key, capacity, meta, ws=None):
        """Called by rasterize_forward under capture, right after the binning was enqueued."""
        if self.key is not None:
            raise RuntimeError("a CaptureProbe serves ONE frame per captured graph")
        self.key, self.capacity = key, int(capacity)
        self.keep.append(ws)
        self.dev[0:2].copy_(meta)
        self.dev[2:3].add_(1)
        self.host.copy_(self.dev, non_blocking=True)

    def wait(self, timeout_s: float = 30.0):
        """Polls until the latest replay has reported; returns (count, needed).  needed != 0: the frame overflowed
        its binning buffer (it rendered as background); the shape's capacity is raised either way."""
        if self.key is None:
            return 0, 0  # the captured callable rendered nothing through rasterize_forward
        import time
        t0 = time.perf_counter()
        while int(self._np[2]) != (self.replays & 0x7FFFFFFF):
            if time.perf_counter() - t0 > timeout_s:
                raise RuntimeError("captured frame never reported its binning result (graph not replayed?)")
            time.sleep(2e-5)  # give the interpreter lock away: other host threads (clock sampling, data loading) run on
        self.count, self.need, self.seen = int(self._np[0]), int(self._np[1]), self.replays
        _raise_capacity(self.key, max(self.count, self.need))
        return self.count, self.need


class capture_probe:
    """Context manager: frames rendered under CUDA-graph capture inside the block report through `probe`."""

    def __init__(self, probe: CaptureProbe):
        self.probe, self.prev = probe, None

    def __enter__(self):
        self.prev, _CAPTURE["probe"] = _CAPTURE["probe"], self.probe
        return self.probe

    def __exit__(self, *exc):
        _CAPTURE["probe"] = self.prev
        return False


# ---- per-(device, stream, P, W, H) workspaces -----------------------------------------------------------
class _Workspace:
    """The caller-owned byte buffers of ONE frame in flight (geom / binning / image / backward scratch),
    recycled through a pool instead of four allocations per call.  A workspace belongs to one stream (work
    on it is stream-ordered, so reuse needs no events) and goes back to the pool when the RasterState that
    checked it out dies -- i.e. after the autograd graph of that frame has been released."""

    __slots__ = ("key", "geom", "image", "binning", "scratch", "meta", "generation")

    def __init__(self, key, dev, P, W, H):
        lib = L.lib()
        self.key = key
        self.geom = torch.empty(lib.dmgs_geom_bytes(P), dtype=torch.uint8, device=dev)
        self.image = torch.empty(lib.dmgs_image_bytes(W, H), dtype=torch.uint8, device=dev)
        self.binning = None
        self.scratch = None
        self.meta = torch.zeros(2, dtype=torch.int32, device=dev)  # {instance count, overflow: count needed}
        self.generation = 0

    def ensure_binning(self, nbytes):
        if self.binning is None or self.binning.numel() < nbytes:
            self.binning = None  # free before growing
            self.binning = torch.empty(int(nbytes), dtype=torch.uint8, device=self.geom.device)
        return self.binning

    def ensure_scratch(self, nbytes):
        if self.scratch is None or self.scratch.numel() < nbytes:
            self.scratch = torch.empty(int(nbytes), dtype=torch.uint8, device=self.geom.device)
        return self.scratch


_POOL: "OrderedDict" = OrderedDict()  # key -> free workspaces, least recently used key first
_POOL_MAX_PER_KEY = 4
_POOL_MAX_KEYS = 16


def _acquire(dev, stream_id, P, W, H) -> _Workspace:
    key = (dev.index, int(stream_id), P, W, H)
    free = _POOL.get(key)
    if free is not None:
        _POOL.move_to_end(key)
    ws = free.pop() if free else _Workspace(key, dev, P, W, H)
    ws.generation += 1
    return ws


def _release(ws: _Workspace):
    free = _POOL.setdefault(ws.key, [])
    if len(free) < _POOL_MAX_PER_KEY:
        free.append(ws)
    while len(_POOL) > _POOL_MAX_KEYS:  # shapes that are no longer rendered give their memory back
        _POOL.popitem(last=False)


def release_workspaces():
    """Drops every pooled workspace (frees the device memory once the allocator's cache is emptied)."""
    _POOL.clear()


def _version_of(t):
    try:
        return t._version
    except Exception:  # inference tensors do not track versions
        return None


def _host_values(tensors):
    """Flat float lists of the small camera tensors (bg, viewmatrix, projmatrix, campos).

    Values are cached per tensor OBJECT (weak reference + in-place version counter), never per
    data pointer: the caching allocator hands the same address to unrelated tensors.  All misses
    of one call share a single device->host copy."""
    out, miss = [None] * len(tensors), []
    for i, t in enumerate(tensors):
        hit = _HOST_CACHE.get(id(t))
        if hit is not None and hit[0]() is t and hit[1] == _version_of(t):
            out[i] = hit[2]
        else:
            miss.append(i)
    if miss:
        if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("CUDA-graph capture: the camera tensors of these settings have not been read to the host "
                               "yet -- render the view once eagerly before capturing it")
        flat = torch.cat([tensors[i].detach().reshape(-1).float() for i in miss]).cpu().tolist()
        o = 0
        for i in miss:
            t = tensors[i]
            n = t.numel()
            vals = flat[o:o + n]
            o += n
            out[i] = vals
            key = id(t)
            ver = _version_of(t)
            if ver is None:
                continue  # inference-mode tensors carry no version counter: never cached
            try:
                _HOST_CACHE[key] = (weakref.ref(t, lambda _r, k=key: _HOST_CACHE.pop(k, None)), ver, vals)
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


def _stream(dev=None):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _check(name, t, shape=None, dtype=torch.float32, device=None):
    """The C ABI takes raw pointers: refuse anything that is not a contiguous CUDA tensor of the expected dtype /
    shape on the expected device (None passes: optional inputs).  `shape` entries of None are free."""
    if t is None:
        return
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: dmgs_b200 needs CUDA tensors; there is no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous (got strides {tuple(t.stride())})")
    if device is not None and t.device != device:
        raise ValueError(f"{name}: on {t.device}, expected {device}")
    if shape is not None:
        if t.dim() != len(shape) or any(e is not None and int(e) != int(g) for e, g in zip(shape, t.shape)):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")


def _check_inputs(P, dev, means3D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp, sh_layout):
    _check("means3D", means3D, (P, 3), device=dev)
    if opacities is not None and opacities.numel() != P:
        raise ValueError(f"opacities: expected {P} elements, got {tuple(opacities.shape)}")
    _check("opacities", opacities, None, device=dev)
    _check("colors_precomp", colors_precomp, (P, 3), device=dev)
    _check("scales", scales, (P, 3), device=dev)
    _check("rotations", rotations, (P, 4), device=dev)
    _check("cov3D_precomp", cov3D_precomp, (P, 6), device=dev)
    _check("shs", shs, (P, None, 3) if sh_layout == 0 else (P, 3, None), device=dev)


class RasterState:
    """Caller-owned state of one forward (the geom / binning / image byte buffers of a pooled workspace)."""

    def __init__(self, prm, ws, num_rendered, radii, capacity=None, probe=None):
        self.prm, self.ws, self.radii, self._probe = prm, ws, radii, probe
        self._gen = ws.generation
        # layout_R sizes the binning layout in every later C call: the instance count on the synchronous
        # path, the buffer capacity on the sync-free path (where the real count stays on the device)
        self.layout_R = num_rendered if capacity is None else capacity
        self._count = num_rendered if capacity is None else None
        self._verified = capacity is None
        self._overflow = 0
        self._captured = None  # CaptureProbe of a frame recorded into a CUDA graph (checked by the graph's owner)

    def __del__(self):
        ws, self.ws = getattr(self, "ws", None), None
        if ws is not None and ws.generation == self._gen and getattr(self, "_captured", None) is None:
            try:
                _release(ws)
            except Exception:
                pass

    def _live(self):
        if self.ws is None or self.ws.generation != self._gen:
            raise RuntimeError("the state buffers of this frame have been recycled")
        return self.ws

    geom = property(lambda self: self._live().geom)
    binning = property(lambda self: self._live().binning)
    image = property(lambda self: self._live().image)

    def verify(self, raise_on_overflow=True) -> bool:
        """Sync-free frames: waits for the frame's (count, overflow) copy and checks it.  False / raises when the
        frame overflowed its binning buffer (the capacity for its shape is raised either way)."""
        if self._captured is not None:  # a frame inside a CUDA graph: its owner polls the CaptureProbe per replay
            if self._captured.count is not None:
                self._count, self._overflow = self._captured.count, self._captured.need
            return not self._overflow
        if not self._verified:
            self._probe.resolve()
            self._count, self._overflow, self._verified = self._probe.count, self._probe.need, True
        if self._overflow and raise_on_overflow:
            raise BinningOverflowError(
                f"sync-free binning: the frame needed {self._overflow} instances but its buffer held {self.layout_R}; "
                "it rendered as background.  The capacity has been raised -- repeat the step "
                "(dmgs_b200.check_async() reports this without raising).")
        return not self._overflow

    @property
    def num_rendered(self) -> int:
        if self._count is None:
            self.verify(raise_on_overflow=False)
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
        with torch.cuda.device(self.geom.device):
            L.check(L.lib().dmgs_sorted_keys(L.ptr(self.geom), L.ptr(self.binning), self.prm.P, R, self.prm.image_width,
                                             self.prm.image_height, L.ptr(keys), _stream(self.geom.device)), "dmgs_sorted_keys")
        return keys[:R]


class BoundMesh(NamedTuple):
    """Mesh source of the fused bind + preprocess path (dmgs_preprocess_forward_bound): Gaussian i = face i // k,
    barycentric row i % k.  verts [V,3] f32, faces [F,3] int64, bc [k,3] f32, g: 1-element f32 tensor
    (tanh(scale_factor) * max_scale) or None."""
    verts: torch.Tensor
    faces: torch.Tensor
    bc: torch.Tensor
    rad_base: float
    thin_z: float
    g: Optional[torch.Tensor]
    adaptive: bool = True


def _check_mesh(m: BoundMesh):
    dev = m.verts.device
    _check("verts", m.verts, (None, 3), device=dev)
    _check("faces", m.faces, (None, 3), dtype=torch.int64, device=dev)
    _check("bc", m.bc, (None, 3), device=dev)
    if m.g is not None:
        _check("g", m.g, (1,), device=dev)
    return dev, int(m.faces.shape[0]), int(m.bc.shape[0])


def rasterize_forward(settings, means3D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp,
                      sh_layout=0, sh_activation=0, stage_hook=None, bound: Optional[BoundMesh] = None, xyz_out=None):
    """Runs the three forward stages; returns (color [3,H,W], radii [P] int32, RasterState).
    Inputs: contiguous float32 CUDA tensors on one device (checked); work is enqueued on that device's current
    stream.  bound: build mean + covariance from the mesh inside preprocess instead of reading means3D /
    cov3D_precomp (which must then be None); xyz_out [F*k,3] optionally receives the means."""
    if bound is not None:
        if means3D is not None or scales is not None or rotations is not None or cov3D_precomp is not None:
            raise ValueError("bound mesh source: means3D / scales / rotations / cov3D_precomp must be None")
        dev, F_, k_ = _check_mesh(bound)
        P = F_ * k_
        _check("xyz_out", xyz_out, (P, 3), device=dev)
    else:
        dev = means3D.device
        P = int(means3D.shape[0])
    if dev.type != "cuda":
        raise RuntimeError("dmgs_b200 rasteriser needs CUDA tensors; there is no CPU path")
    lib = L.lib()
    H, W = int(settings.image_height), int(settings.image_width)
    if bound is None:
        _check_inputs(P, dev, means3D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp, sh_layout)
    else:
        _check_inputs(P, dev, None, opacities, shs, colors_precomp, None, None, None, sh_layout)
    M = 0
    if shs is not None:
        M = int(shs.shape[1] if sh_layout == 0 else shs.shape[2])
        if M < (int(settings.sh_degree) + 1) ** 2:
            raise ValueError(f"shs holds {M} coefficients per channel, sh_degree {settings.sh_degree} needs "
                             f"{(int(settings.sh_degree) + 1) ** 2}")
    prm = make_params(settings, P, M, sh_layout, sh_activation)
    with torch.cuda.device(dev):
        cur = torch.cuda.current_stream(dev)
        stream = C.c_void_p(cur.cuda_stream)
        radii = torch.zeros(P, dtype=torch.int32, device=dev)
        color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
        ws = _acquire(dev, cur.cuda_stream, P, W, H)
        nr, flag = ws.meta[0:1], ws.meta[1:2]
        key = (dev.index, P, W, H)
        cap = None
        if _ASYNC["on"] and key not in _ASYNC["no_async"]:
            cap = _ASYNC["capacity"].get(key)
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing and cap is None:
            raise RuntimeError("CUDA-graph capture needs sync-free binning with a known capacity: "
                               "configure(async_binning=True) and render this shape once eagerly first")
        if bound is None:
            L.check(lib.dmgs_preprocess_forward(C.byref(prm), L.ptr(means3D), L.ptr(scales), L.ptr(rotations),
                                                L.ptr(cov3D_precomp), L.ptr(opacities), L.ptr(shs), L.ptr(colors_precomp),
                                                L.ptr(radii), L.ptr(ws.geom), L.ptr(nr), stream), "dmgs_preprocess_forward")
        else:
            L.check(lib.dmgs_preprocess_forward_bound(C.byref(prm), F_, k_, L.ptr(bound.verts), L.ptr(bound.faces),
                                                      L.ptr(bound.bc), float(bound.rad_base), float(bound.thin_z),
                                                      L.ptr(bound.g), int(bool(bound.adaptive)), L.ptr(opacities), L.ptr(shs),
                                                      L.ptr(colors_precomp), L.ptr(radii), L.ptr(ws.geom), L.ptr(nr),
                                                      L.ptr(xyz_out), stream), "dmgs_preprocess_forward_bound")
        if stage_hook is not None:
            stage_hook("preprocess_sort_scan")
        state = None
        if cap is not None:
            # sync-free: buffer sized from earlier frames of this shape; the count stays on the device and is
            # copied, with the overflow flag, to pinned memory behind an event (RasterState.verify)
            R = int(cap)
            binning = ws.ensure_binning(lib.dmgs_binning_bytes(P, R, W, H))
            rc = lib.dmgs_bin_forward_async(C.byref(prm), L.ptr(ws.geom), R, L.ptr(binning), L.ptr(flag), stream)
            if rc == -9:  # image too large for the placement path: this shape stays synchronous
                _ASYNC["no_async"].add(key)
                _ASYNC["capacity"].pop(key, None)
                cap = None
            elif capturing:
                L.check(rc, "dmgs_bin_forward_async")
                if _CAPTURE["probe"] is None:
                    raise RuntimeError("rasterize_forward under CUDA-graph capture: wrap the capture in "
                                       "dmgs_b200.rasterizer.capture_probe(CaptureProbe(device)) "
                                       "(multiview.ViewStreams.capture does)")
                _CAPTURE["probe"].record(key, R, ws.meta, ws)
                state = RasterState(prm, ws, None, radii, capacity=R)
                state._captured = _CAPTURE["probe"]
            else:
                L.check(rc, "dmgs_bin_forward_async")
                probe = _Probe(key, R, ws.meta, cur)
                state = RasterState(prm, ws, None, radii, capacity=R, probe=probe)
                pend = _ASYNC["pending"].setdefault(dev.index, [])
                if len(pend) >= _MAX_PENDING:  # nobody polls check_async(): resolve the oldest frames ourselves
                    old, pend[:] = pend[:_MAX_PENDING // 2], pend[_MAX_PENDING // 2:]
                    if not all([pr_.resolve() for pr_ in old]):
                        raise BinningOverflowError("sync-free binning: earlier frames overflowed their binning buffer and "
                                                   "nobody called dmgs_b200.check_async(); capacities have been raised")
                pend.append(probe)
        if cap is None:
            R = int(nr.item())  # the one host read-back of the forward (upstream does the same after its scan)
            if _ASYNC["on"] and key not in _ASYNC["no_async"]:
                _raise_capacity(key, R)
            binning = ws.ensure_binning(lib.dmgs_binning_bytes(P, R, W, H))
            L.check(lib.dmgs_bin_forward(C.byref(prm), L.ptr(ws.geom), R, L.ptr(binning), stream), "dmgs_bin_forward")
            state = RasterState(prm, ws, R, radii)
        if stage_hook is not None:
            stage_hook("binning")
        L.check(lib.dmgs_blend_forward(C.byref(prm), L.ptr(ws.geom), L.ptr(binning), R, L.ptr(color), L.ptr(ws.image),
                                       stream), "dmgs_blend_forward")
        if stage_hook is not None:
            stage_hook("blend_fwd")
    return color, radii, state


def rasterize_backward_bound(state: RasterState, grad_color, bound: BoundMesh, shs, want_colors_precomp, dverts, dg,
                             verify=True):
    """Backward of a frame rendered with `rasterize_forward(..., bound=mesh)`: blend backward, then the per-Gaussian
    backward with the binding adjoint fused in (dmgs_preprocess_backward_bound).  dverts [V,3] and dg [1] (or None)
    are ADDED to (the caller zeroes them).  Returns (dL/dmeans2D [P,3], dL/dshs | None, dL/dcolors_precomp | None,
    dL/dopacities [P,1])."""
    lib = L.lib()
    prm, P = state.prm, state.prm.P
    dev, F_, k_ = _check_mesh(bound)
    _check("grad_color", grad_color, (3, prm.image_height, prm.image_width), device=dev)
    _check("shs", shs, (P, None, 3) if prm.sh_layout == 0 else (P, 3, None), device=dev)
    _check("dverts", dverts, tuple(bound.verts.shape), device=dev)
    _check("dg", dg, (1,), device=dev)
    if verify:
        state.verify()
    ws = state._live()
    z = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    g_means2D, g_op = z(P, 3), z(P, 1)
    g_col = z(P, 3) if want_colors_precomp else None
    g_shs = torch.empty_like(shs) if shs is not None else None
    with torch.cuda.device(dev):
        stream = _stream(dev)
        scratch = ws.ensure_scratch(lib.dmgs_backward_scratch_bytes(P))
        L.check(lib.dmgs_blend_backward(C.byref(prm), L.ptr(ws.geom), L.ptr(ws.binning), L.ptr(ws.image),
                                        state.layout_R, L.ptr(grad_color), L.ptr(scratch), stream), "dmgs_blend_backward")
        L.check(lib.dmgs_preprocess_backward_bound(C.byref(prm), F_, k_, L.ptr(bound.verts), L.ptr(bound.faces),
                                                   L.ptr(bound.bc), float(bound.rad_base), float(bound.thin_z),
                                                   L.ptr(bound.g), int(bool(bound.adaptive)), L.ptr(shs), L.ptr(state.radii),
                                                   L.ptr(ws.geom), L.ptr(scratch), L.ptr(dverts), L.ptr(dg), L.ptr(g_means2D),
                                                   L.ptr(g_op), L.ptr(g_col), L.ptr(g_shs), 0, stream),
                "dmgs_preprocess_backward_bound")
    return g_means2D, g_shs, g_col, g_op


def rasterize_backward(state: RasterState, grad_color, means3D, shs, scales, rotations, cov3D_precomp,
                       want_colors_precomp, stage_hook=None, accumulate_into=None, sh_record=None, verify=True):
    """accumulate_into: optional dict of preallocated gradient tensors (keys means3D, means2D, opacities,
    colors_precomp, shs, scales, rotations, cov3D_precomp) that the gradients are ADDED to.
    sh_record: with accumulate_into, a float32 [P,4] tensor that receives this view's deferred SH gradient
    record {g.r, g.g, g.b, seen} instead of the [P,16,3] rows being read-modify-written
    (dmgs_preprocess_backward accumulate = 2; `sh_grad_expand` forms the rows once per step).
    verify: sync-free frames are checked for a binning overflow first (BinningOverflowError); a training step that
    polls `check_async()` itself once per step passes verify=False to keep the host running ahead."""
    lib = L.lib()
    prm, P, dev = state.prm, state.prm.P, means3D.device
    H, W = prm.image_height, prm.image_width
    sh_layout = prm.sh_layout
    _check("grad_color", grad_color, (3, H, W), device=dev)
    _check_inputs(P, dev, means3D, None, shs, None, scales, rotations, cov3D_precomp, sh_layout)
    if verify:
        state.verify()
    ws = state._live()
    z = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    if accumulate_into is None:
        g_means3D, g_means2D, g_op = z(P, 3), z(P, 3), z(P, 1)
        g_col = z(P, 3) if want_colors_precomp else None
        g_shs = torch.empty_like(shs) if shs is not None else None
        g_scales = z(P, 3) if scales is not None else None
        g_rots = z(P, 4) if rotations is not None else None
        g_cov = z(P, 6) if cov3D_precomp is not None else None
    else:
        a = accumulate_into
        g_means3D, g_means2D, g_op = a["means3D"], a["means2D"], a["opacities"]
        g_col = a.get("colors_precomp") if want_colors_precomp else None
        g_shs = a.get("shs") if shs is not None else None
        g_scales = a.get("scales") if scales is not None else None
        g_rots = a.get("rotations") if rotations is not None else None
        g_cov = a.get("cov3D_precomp") if cov3D_precomp is not None else None
        _check("accumulate_into[means3D]", g_means3D, (P, 3), device=dev)
        _check("accumulate_into[means2D]", g_means2D, (P, 3), device=dev)
        _check("accumulate_into[opacities]", g_op, None, device=dev)
        if g_op.numel() != P:
            raise ValueError("accumulate_into[opacities] must hold P elements")
        _check("accumulate_into[colors_precomp]", g_col, (P, 3), device=dev)
        _check("accumulate_into[scales]", g_scales, (P, 3), device=dev)
        _check("accumulate_into[rotations]", g_rots, (P, 4), device=dev)
        _check("accumulate_into[cov3D_precomp]", g_cov, (P, 6), device=dev)
        if shs is not None and g_shs is not None:
            _check("accumulate_into[shs]", g_shs, tuple(shs.shape), device=dev)
        if sh_record is not None and shs is not None:
            _check("sh_record", sh_record, (P, 4), device=dev)
            g_shs = sh_record
    with torch.cuda.device(dev):
        stream = _stream(dev)
        scratch = ws.ensure_scratch(lib.dmgs_backward_scratch_bytes(P))
        L.check(lib.dmgs_blend_backward(C.byref(prm), L.ptr(ws.geom), L.ptr(ws.binning), L.ptr(ws.image),
                                        state.layout_R, L.ptr(grad_color), L.ptr(scratch), stream),
                "dmgs_blend_backward")
        if stage_hook is not None:
            stage_hook("blend_bwd")
        L.check(lib.dmgs_preprocess_backward(C.byref(prm), L.ptr(means3D), L.ptr(scales), L.ptr(rotations),
                                             L.ptr(cov3D_precomp), L.ptr(shs), L.ptr(state.radii), L.ptr(ws.geom),
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
    if records.dim() != 3 or records.shape[2] != 4 or records.stride(2) != 1 or records.stride(1) != 4:
        raise ValueError(f"records: expected [V,P,4] with contiguous [P,4] views, got {tuple(records.shape)} / "
                         f"{tuple(records.stride())}")
    V, P = int(records.shape[0]), int(records.shape[1])
    dev = records.device
    _check("means3D", means3D, (P, 3), device=dev)
    _check("shs", shs, (P, None, 3) if sh_layout == 0 else (P, 3, None), device=dev)
    _check("out", out, tuple(shs.shape), device=dev)
    _check("out_means3D", out_means3D, (P, 3), device=dev)
    if records.dtype != torch.float32 or not records.is_cuda:
        raise TypeError("records: expected a float32 CUDA tensor")
    M = int(out.shape[1] if sh_layout == 0 else out.shape[2])
    if M < (int(sh_degree) + 1) ** 2:
        raise ValueError(f"out holds {M} coefficients per channel, sh_degree {sh_degree} needs {(int(sh_degree) + 1) ** 2}")
    cams = []
    for c in camera_centers:
        cams.extend(_host_values([c])[0] if isinstance(c, torch.Tensor) else [float(x) for x in c])
    if len(cams) != 3 * V:
        raise ValueError("one camera centre per view")
    arr = (C.c_float * (3 * V))(*cams)
    stride = int(records.stride(0)) if V > 1 else P * 4
    with torch.cuda.device(dev):
        L.check(L.lib().dmgs_sh_grad_expand(P, int(sh_degree), M, int(sh_layout), V, arr, L.ptr(means3D), L.ptr(shs),
                                            L.ptr(records), stride, L.ptr(out), L.ptr(out_means3D), int(bool(accumulate)),
                                            _stream(dev)), "dmgs_sh_grad_expand")
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
        # the caller's own tensors, for configure(accumulate_grad_in_place=True)
        ctx.leaves = (means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp) if _FUSED_ACC["on"] else None
        if holder is not None:
            holder.last = state
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, _grad_radii):
        if ctx.state is None:
            return (None,) * 12
        means3D, sh, scales, rotations, cov = ctx.saved_tensors
        leaves, ctx.leaves = ctx.leaves, None
        if leaves is not None and _FUSED_ACC["on"]:
            fused = _fused_accumulate(ctx, leaves, grad_color, means3D, sh, scales, rotations, cov)
            if fused is not None:
                return fused
        g = rasterize_backward(ctx.state, grad_color.contiguous().float(), means3D, sh, scales, rotations, cov,
                               ctx.has_col)
        g_means3D, g_means2D, g_shs, g_col, g_op, g_scales, g_rots, g_cov = g
        return (g_means3D, g_means2D, g_shs, g_col, g_op, g_scales, g_rots, g_cov, None, None, None, None)


def _fused_accumulate(ctx, leaves, grad_color, means3D, sh, scales, rotations, cov):
    """configure(accumulate_grad_in_place=True): gradients of inputs whose `.grad` exists are added into it by the
    kernels (autograd gets None for them); the others get a fresh zero-initialised tensor that the kernels add into.
    Returns None when nothing would be gained (no input has a usable `.grad`)."""
    names = ("means3D", "means2D", "shs", "colors_precomp", "opacities", "scales", "rotations", "cov3D_precomp")
    P = ctx.state.prm.P
    acc, ret, hit = {}, [None] * 8, False
    for k, (name, t) in enumerate(zip(names, leaves)):
        if t is None or not ctx.needs_input_grad[k]:
            continue
        g = t.grad if t.is_leaf else None
        ok = (g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.shape == t.shape and g.is_cuda
              and not getattr(t, "_backward_hooks", None))
        if ok:
            acc[name], hit = g, True
        else:
            acc[name] = ret[k] = torch.zeros(t.shape, dtype=torch.float32, device=means3D.device)
    if not hit:
        return None
    # the C entry point always writes these three: inputs that need no gradient get a scratch tensor
    for name, shape in (("means3D", (P, 3)), ("means2D", (P, 3)), ("opacities", (P, 1))):
        if name not in acc:
            acc[name] = torch.zeros(shape, dtype=torch.float32, device=means3D.device)
    for name, t in (("shs", sh), ("scales", scales), ("rotations", rotations), ("cov3D_precomp", cov)):
        if t is not None and name not in acc:
            acc[name] = torch.zeros(t.shape, dtype=torch.float32, device=means3D.device)
    if ctx.has_col and "colors_precomp" not in acc:
        acc["colors_precomp"] = torch.zeros(P, 3, dtype=torch.float32, device=means3D.device)
    rasterize_backward(ctx.state, grad_color.contiguous().float(), means3D, sh, scales, rotations, cov, ctx.has_col,
                       accumulate_into=acc)
    return tuple(ret) + (None, None, None, None)


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
                with torch.cuda.device(positions.device):
                    L.check(L.lib().dmgs_mark_visible(P, L.ptr(pos), L.ptr(view), L.ptr(proj), L.ptr(vis),
                                                      _stream(positions.device)), "dmgs_mark_visible")
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
