"""View-partitioned multi-GPU training step: replicated Gaussians, the camera batch of a step split
over the ranks of one box, one all-reduce of the flat per-Gaussian gradient buffer per step.

The reference is single-view, single-GPU (train_geo_stage2.py:91-93 pops ONE camera per iteration,
utils/general_utils.py:133 pins cuda:0), so this module is new capability,
not a mirror.  One process per GPU (torchrun); `torch.distributed` is the plumbing (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  There is no data-path collective other than the one
gradient all-reduce: every view's render is independent given identical Gaussians.

    buf  = FlatGradBuffer(P, RASTER_WIDTHS_SH, device)
    vp   = ViewParallel()                      # reads the default process group
    loss = vp.step(n_views, lambda v, acc: render_and_accumulate(v, acc), buf)
    # buf.views["means3D"] ... now hold the gradient summed over ALL views of the step, on every rank
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

# floats per Gaussian of every gradient the rasteriser returns, per input mode
RASTER_WIDTHS_SH = {"means3D": (3,), "means2D": (3,), "opacities": (1,), "scales": (3,), "rotations": (4,),
                    "shs": (16, 3)}
RASTER_WIDTHS_PRECOMP = {"means3D": (3,), "means2D": (3,), "opacities": (1,), "colors_precomp": (3,),
                         "cov3D_precomp": (6,)}
RASTER_WIDTHS_STAGE2_FUSED = {"means3D": (3,), "means2D": (3,), "opacities": (1,), "cov3D_precomp": (6,),
                              "shs": (3, 16)}


def partition_views(n_views: int, world: int, rank: int) -> List[int]:
    """Views of a step rendered by `rank`: v = rank (mod world) -- round-robin keeps ranks within one
    view of each other for any n_views."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_views, world))


class FlatGradBuffer:
    """One contiguous fp32 buffer holding every per-Gaussian gradient, field-major, with named
    [P, ...] views.  The views are what `rasterize_backward(..., accumulate_into=views)` adds to and
    the flat tensor is the single all-reduce payload of a step."""

    def __init__(self, P: int, widths: Dict[str, Tuple[int, ...]], device, dtype=torch.float32, allocate=None):
        """allocate(numel) -> 1-D fp32 tensor: where the flat buffer lives (default: torch.zeros; the
        peer-memory all-reduce passes a symmetric-memory allocator)."""
        self.P, self.widths = int(P), dict(widths)
        n = 0
        self.offsets: Dict[str, Tuple[int, int]] = {}
        for name, tail in self.widths.items():
            w = 1
            for t in tail:
                w *= int(t)
            # 16-byte aligned fields: the SH rows are written with bulk (TMA) reductions
            n = (n + 3) // 4 * 4
            self.offsets[name] = (n, self.P * w)
            n += self.P * w
        n = (n + 3) // 4 * 4  # whole 16-byte groups: vector kernels (Adam, peer all-reduce) need no tail
        self.flat = torch.zeros(n, dtype=dtype, device=device) if allocate is None else allocate(n)
        self.views = {name: self.flat[o:o + m].view(self.P, *self.widths[name]) for name, (o, m) in self.offsets.items()}

    def zero_(self):
        self.flat.zero_()
        return self

    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()


class ViewParallel:
    """Runs the local views of a step and all-reduces the flat gradient buffer."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0

    def local_views(self, n_views: int) -> List[int]:
        return partition_views(n_views, self.world, self.rank)

    def step(self, n_views: int, render_view: Callable[[int, Dict[str, torch.Tensor]], Optional[torch.Tensor]],
             buf: FlatGradBuffer, average: bool = False, extra: Sequence[torch.Tensor] = ()) -> Optional[torch.Tensor]:
        """render_view(v, buf.views) must ADD view v's gradients into the views and may return the
        view's loss (a 0-d tensor).  Returns the loss summed (or averaged) over all views of the
        step.  `extra` tensors (e.g. dL/dverts of a mesh-bound model) are all-reduced as well."""
        buf.zero_()
        loss = None
        for v in self.local_views(n_views):
            l = render_view(v, buf.views)
            if l is not None:
                loss = l.detach().clone() if loss is None else loss + l.detach()
        if self.world > 1:
            dist.all_reduce(buf.flat, op=dist.ReduceOp.SUM, group=self.group)
            for t in extra:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            if loss is None:
                loss = torch.zeros((), dtype=buf.flat.dtype, device=buf.flat.device)
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        if average and n_views > 0:
            buf.flat.div_(n_views)
            for t in extra:
                t.div_(n_views)
            if loss is not None:
                loss = loss / n_views
        return loss


class SymmetricFlat:
    """A flat fp32 buffer in torch symmetric memory, mapped into every rank of `group` (CUDA VMM peer mappings plus,
    on NVSwitch systems, the multicast mapping).  Construction is a collective (rendezvous).  `allocate` has the
    signature FlatGradBuffer's `allocate=` expects, so gradients AND parameters can live in such buffers."""

    def __init__(self, numel: int, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.flat = symm.empty(int(numel), dtype=torch.float32, device=device)
        self.flat.zero_()
        self.hdl = symm.rendezvous(self.flat, self.group)
        self.world, self.rank = int(self.hdl.world_size), int(self.hdl.rank)
        self.peers = [self.flat if k == self.rank else self.hdl.get_buffer(k, (int(numel),), torch.float32)
                      for k in range(self.world)]  # keep the mappings alive
        self.ptrs = (C.c_void_p * self.world)(*[p.data_ptr() for p in self.peers])
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        if mc:
            base = int(self.hdl.buffer_ptrs[self.rank])
            mc += self.flat.data_ptr() - base  # the tensor's offset inside the symmetric allocation
        self.multicast_ptr = mc

    @classmethod
    def allocator(cls, device, group=None):
        """-> (allocate(numel) -> tensor, holder list that receives the SymmetricFlat)."""
        holder: List["SymmetricFlat"] = []

        def allocate(numel):
            holder.append(cls(numel, device, group))
            return holder[-1].flat
        return allocate, holder


class PeerAllReduce:
    """Sum of a flat fp32 buffer over the ranks of one box through NVLink peer memory, in place
    (libdmgs_raster.so: dmgs_allreduce_peer) -- the gradient exchange of the view-partitioned step
    without NCCL: rank r reduces slice r of every rank's buffer and writes the (scaled) sum back to
    slice r of every rank's buffer; on NVSwitch systems with multicast the addition happens inside the
    switch (multimem.ld_reduce / multimem.st).

    The buffer lives in torch symmetric memory (`allocate` hands it to FlatGradBuffer); construction is
    a collective (rendezvous).  `all_reduce_()` enqueues barrier -> kernel -> barrier on the current
    stream; nothing synchronises with the host."""

    MULTICAST_MIN_WORLD = 5  # measured on B200 x8 (profiles/r1_s6_allreduce.jsonl): P2P wins at 2 and 4 ranks, multimem at 8

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self._symm, self.device = symm, device
        self.group = group if group is not None else dist.group.WORLD
        self.flat = self.hdl = None

    def allocate(self, numel: int) -> torch.Tensor:
        if self.flat is not None:
            raise RuntimeError("PeerAllReduce serves one buffer")
        sym = SymmetricFlat(numel, self.device, self.group)
        self._sym = sym
        self.flat, self.hdl, self.world, self.rank = sym.flat, sym.hdl, sym.world, sym.rank
        self._peers, self._ptrs, self.multicast_ptr = sym.peers, sym.ptrs, sym.multicast_ptr
        self.ptrs = sym.ptrs
        return self.flat

    def all_reduce_(self, scale: float = 1.0, use_multicast: Optional[bool] = None):
        """use_multicast: None = automatic (in-switch reduction from MULTICAST_MIN_WORLD ranks up: with few
        ranks the peer loads/stores move as few bytes per link and were measured faster, see
        profiles/), True / False force a path."""
        import ctypes as C
        from . import _lib as L
        if use_multicast is None:
            use_multicast = self.world >= self.MULTICAST_MIN_WORLD
        mc = C.c_void_p(self.multicast_ptr) if (use_multicast and self.multicast_ptr) else None
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self.hdl.barrier(channel=0)  # every rank's accumulator is complete
        L.check(L.lib().dmgs_allreduce_peer(self.flat.numel(), self.world, self.rank, self._ptrs, mc, float(scale), st),
                "dmgs_allreduce_peer")
        self.hdl.barrier(channel=1)  # every slice has been delivered to every rank
        return self.flat


class _CapturedView:
    """One view of a step as a CUDA graph (ViewStreams.capture)."""

    __slots__ = ("graph", "probe", "keep", "stream", "sh_records", "kernel_launches")

    def __init__(self, graph, probe, keep, stream, sh_records):
        self.graph, self.probe, self.keep, self.stream, self.sh_records = graph, probe, keep, stream, list(sh_records)
        self.kernel_launches = 0


class ViewStreams:
    """Renders the local views of a step on `n` CUDA streams, round-robin, all of them accumulating into ONE
    FlatGradBuffer; `finish()` joins the streams (and, in deferred-SH mode, forms the SH rows).

    Why: the stages of one view alternate between latency-bound kernels (depth sort, tile placement:
    < 50 % issue utilisation, little HBM traffic) and issue-bound ones (the blend kernels: ~75 % issue
    utilisation, ~1 % of the HBM bandwidth).  Views of a step are independent given identical
    Gaussians, so several views in flight let the SMs fill one view's stalls with the other's work.  The
    accumulate modes of the per-Gaussian backward add with global reductions (red.global.add / TMA reduce-add,
    resolved in L2), so concurrent views can share the buffer: no per-stream accumulators, no sums at the end."""

    def __init__(self, P: int, widths: Dict[str, Tuple[int, ...]], device, n: int = 2, peer_group=None,
                 deferred_sh_views: int = 0, sh_key: str = "shs"):
        """deferred_sh_views: > 0 enables the deferred SH gradient for up to that many local views per step:
        every view records only its 16-byte {dL/dcolour, seen} per Gaussian (`sh_record(i, campos)` hands the
        view its record array) and `finish(means3D, shs, sh_degree)` forms the [P,16,3] gradient rows once from all
        records -- V*16 B + one row per Gaussian and step instead of V read-modify-writes of the 192-byte row.
        peer_group: a process group of ranks on ONE box -> the summed buffer lives in symmetric memory
        and `all_reduce_()` exchanges it over NVLink peer memory (falls back to NCCL if symmetric memory
        cannot be set up); None -> `all_reduce_()` is torch.distributed.all_reduce."""
        if n < 1:
            raise ValueError("need at least one stream")
        self.peer, self.peer_error, self.group = None, None, peer_group
        alloc = None
        if peer_group is not None and dist.get_world_size(peer_group) > 1:
            try:
                self.peer = PeerAllReduce(device, peer_group)
                probe = self.peer.allocate  # rendezvous happens on first allocation (collective)
                alloc = probe
            except Exception as e:  # symmetric memory unavailable: NCCL path
                self.peer, self.peer_error = None, f"{type(e).__name__}: {e}"
        try:
            first = FlatGradBuffer(P, widths, device, allocate=alloc)
        except Exception as e:
            if alloc is None:
                raise
            self.peer, self.peer_error = None, f"{type(e).__name__}: {e}"
            first = FlatGradBuffer(P, widths, device)
        self.bufs = [first]
        self.sh_key, self.P = sh_key, int(P)
        self.records = (torch.zeros(int(deferred_sh_views), int(P), 4, dtype=torch.float32, device=device)
                        if deferred_sh_views > 0 and sh_key in widths else None)
        self._campos, self._used = {}, 0
        if self.records is not None:
            # the SH field is last in the flat layout: the part before it is what begin() zeroes and finish() sums
            off, _ = first.offsets[sh_key]
            if any(o > off for o, _ in first.offsets.values()):
                raise ValueError("deferred SH needs the SH field last in the flat buffer")
            self._dense = off
        # (descending stream priorities were measured and rejected: H0, 4 streams, 1694 -> 1631 frames/s)
        self.streams = [torch.cuda.Stream(device=device) for _ in range(n)] if n > 1 else [None]
        self.device = device
        self._graphs: Dict[int, "_CapturedView"] = {}
        self._capture_log = None
        self.replayed_kernel_launches = 0  # kernels of libdmgs_raster.so launched through graph replays
        # CTAs per SM of the persistent blend kernels while this object's views are being enqueued: with several
        # views in flight 6 leaves a quarter of every SM's registers to the other views' latency-bound stages
        # (H0: 1664 -> 1700 frames/s); a single stream keeps the library default of 8
        import os
        self.blend_residency = int(os.environ.get("DMGS_VIEW_BLEND_RESIDENCY", "6" if n > 1 else "8"))
        # shared memory per SM of the tile-placement kernels: at the library's 200 KB nothing fits beside them; at 128 KB
        # six blend CTAs of another view do (H0: 1698 -> 1739 frames/s, the placement itself 4 % slower)
        self.place_smem_kb = int(os.environ.get("DMGS_VIEW_PLACE_SMEM_KB", "128" if n > 1 else "200"))

    @property
    def buf(self) -> FlatGradBuffer:
        return self.bufs[0]

    @staticmethod
    def _set_residency(k: int, place_kb: int):
        from . import _lib as L
        if torch.cuda.is_available():
            L.check(L.lib().dmgs_set_blend_residency(int(k), int(k)), "dmgs_set_blend_residency")
            L.check(L.lib().dmgs_set_place_smem_kb(int(place_kb)), "dmgs_set_place_smem_kb")

    def begin(self):
        """Zeroes the accumulator; the side streams start after everything queued on the current stream."""
        cur = torch.cuda.current_stream(self.device)
        self._campos, self._used = {}, 0
        self._set_residency(self.blend_residency, self.place_smem_kb)
        if self.records is not None:
            self.bufs[0].flat[:self._dense].zero_()
        else:
            self.bufs[0].zero_()
        for st in self.streams:
            if st is not None:
                st.wait_stream(cur)

    def run(self, i: int, fn: Callable[[Dict[str, torch.Tensor]], object]):
        """Calls fn(acc_views) for the i-th local view on stream i mod n."""
        k = i % len(self.streams)
        if self.streams[k] is None:
            return fn(self.bufs[0].views)
        with torch.cuda.stream(self.streams[k]):
            return fn(self.bufs[0].views)

    def sh_record(self, i: int, campos) -> Optional[torch.Tensor]:
        """Record array [P,4] of the i-th local view of the step (deferred SH mode; None otherwise); campos is
        the view's camera centre (settings.campos)."""
        if self.records is None:
            return None
        if not 0 <= i < self.records.shape[0]:
            raise IndexError(f"view {i}: ViewStreams was built for {self.records.shape[0]} deferred-SH views per step")
        self._campos[i] = campos
        self._used = max(self._used, i + 1)
        if self._capture_log is not None:  # a captured view takes the same record on every replay
            self._capture_log.append((i, campos))
        return self.records[i]

    # ---- views as CUDA graphs --------------------------------------------------------------------------
    def capture(self, i: int, fn: Callable[[Dict[str, torch.Tensor]], object]):
        """Captures what fn(acc_views) ENQUEUES for the i-th local view (forward, loss gradient, backward with
        `accumulate_into`, ...) into a CUDA graph on stream i mod n; fn is not executed on the device.  `replay(i)`
        then launches the whole view with one graph launch -- the host needs ~10 us per view instead of ~250 us for
        the ~25 launches and their glue, so every stream has its work at the very start of the step.

        Requirements: sync-free binning (`dmgs_b200.configure(async_binning=True)`) and one eager frame of this shape
        and these settings beforehand (capacity known, camera tensors cached on the host); every tensor fn reads or
        writes must keep its address (parameters updated IN PLACE, the accumulator of this object, ...).  The frame's
        workspace stays with the graph.  After the step `poll_captured()` must be called: it reports frames that
        overflowed their binning buffer (capacity raised, graph dropped -> capture again, repeat the step)."""
        from . import rasterizer as R
        k = i % len(self.streams)
        if self.streams[k] is None:  # capture needs a non-default stream; begin() / finish() fork and join it
            self.streams[k] = torch.cuda.Stream(device=self.device)
        st = self.streams[k]
        from . import _lib as L
        probe = R.CaptureProbe(self.device)
        graph = torch.cuda.CUDAGraph()
        self._capture_log = []
        n0 = int(L.lib().dmgs_launch_count())
        try:
            with R.capture_probe(probe), torch.cuda.graph(graph, stream=st, capture_error_mode="thread_local"):
                keep = fn(self.bufs[0].views)
            self._graphs[i] = _CapturedView(graph, probe, keep, st, self._capture_log)
            # kernels of this library inside the graph: a replay launches them again without passing through the
            # library's launch counter (dmgs_launch_count), so replay() accounts for them here
            self._graphs[i].kernel_launches = int(L.lib().dmgs_launch_count()) - n0
        finally:
            self._capture_log = None
        return self._graphs[i]

    def captured(self, i: int) -> bool:
        return i in self._graphs

    def replay(self, i: int):
        """Launches the captured graph of the i-th local view on its stream (between begin() and finish())."""
        cv = self._graphs[i]
        for j, campos in cv.sh_records:
            self._campos[j] = campos
            self._used = max(self._used, j + 1)
        with torch.cuda.stream(cv.stream):
            cv.graph.replay()
        cv.probe.replays += 1
        self.replayed_kernel_launches += cv.kernel_launches
        return cv.keep

    def poll_captured(self) -> bool:
        """True if every captured view replayed since the last call fitted its binning buffer.  Polls pinned host
        memory the graphs write right after their binning (no CUDA synchronisation; returns once the LAST view's
        binning is done, long before the step ends).  Overflowed views lose their graph: capture again and repeat
        the step (their frames rendered as background with zero gradients)."""
        ok = True
        for i in list(self._graphs):
            cv = self._graphs[i]
            if cv.probe.seen == cv.probe.replays:
                continue
            _, need = cv.probe.wait()
            if need:
                ok = False
                del self._graphs[i]
        return ok

    def drop_graphs(self):
        self._graphs.clear()

    def finish(self, means3D: Optional[torch.Tensor] = None, shs: Optional[torch.Tensor] = None, sh_degree: int = 3,
               sh_layout: int = 0, means3D_key: str = "means3D") -> FlatGradBuffer:
        """Joins the streams on the current stream and returns the buffer holding the sum of all views
        (deferred SH mode: means3D / shs / sh_degree / sh_layout of the step are needed to form the SH rows and
        the view-direction term of the means3D gradient)."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            if st is not None:
                cur.wait_stream(st)
        self._set_residency(8, 200)
        if self.records is None:
            return self.bufs[0]
        if means3D is None or shs is None:
            raise ValueError("deferred SH mode: finish(means3D, shs, sh_degree) forms the SH gradient rows")
        if sorted(self._campos) != list(range(self._used)):
            raise RuntimeError("deferred SH mode: every view 0..n-1 of the step must have taken its sh_record")
        from .rasterizer import sh_grad_expand
        out = self.bufs[0].views[self.sh_key]
        V, step = self._used, 32
        if V == 0:
            out.zero_()
        for v0 in range(0, V, step):  # DMGS_MAX_STEP_VIEWS views per launch
            v1 = min(V, v0 + step)
            sh_grad_expand(self.records[v0:v1], [self._campos[v] for v in range(v0, v1)], means3D, shs, sh_degree, out,
                           self.bufs[0].views[means3D_key], sh_layout=sh_layout, accumulate=v0 > 0)
        return self.bufs[0]

    def all_reduce_(self, scale: float = 1.0) -> FlatGradBuffer:
        """Sums the (already stream-summed) buffer over the ranks, in place, on the current stream."""
        if self.peer is not None:
            self.peer.all_reduce_(scale)
        elif dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.bufs[0].flat, op=dist.ReduceOp.SUM, group=self.group)
            if scale != 1.0:
                self.bufs[0].flat.mul_(scale)
        elif scale != 1.0:
            self.bufs[0].flat.mul_(scale)
        return self.bufs[0]


def accumulate_view(settings, inputs: Dict[str, Optional[torch.Tensor]], image_grad: Callable, acc: Dict[str, torch.Tensor],
                    sh_layout: int = 0, sh_activation: int = 0, sh_record=None):
    """One view forward + backward on the CUDA path with gradients ADDED into `acc`.

    inputs: means3D, opacities and the optional shs / colors_precomp / scales / rotations / cov3D_precomp.
    sh_record: ViewStreams.sh_record(i, settings.campos) in deferred-SH mode (see ViewStreams).
    image_grad(image) -> (loss 0-d tensor or None, dL/dimage [3,H,W]).  Returns (loss, image, radii)."""
    from .rasterizer import rasterize_backward, rasterize_forward
    g = inputs.get
    color, radii, state = rasterize_forward(settings, inputs["means3D"], inputs["opacities"], g("shs"),
                                            g("colors_precomp"), g("scales"), g("rotations"), g("cov3D_precomp"),
                                            sh_layout, sh_activation)
    loss, dL = image_grad(color)
    rasterize_backward(state, dL.contiguous(), inputs["means3D"], g("shs"), g("scales"), g("rotations"),
                       g("cov3D_precomp"), g("colors_precomp") is not None, accumulate_into=acc, sh_record=sh_record,
                       verify=False)  # sync-free binning: the step polls dmgs_b200.check_async() once, after its views
    return loss, color, radii


class ShardedPeerAdam:
    """Adam over REPLICATED parameters with the gradient exchange fused in (libdmgs_raster.so:
    dmgs_adam_exchange_peer) -- the tail of a view-partitioned training step in one kernel per rank instead of
    "all-reduce, then the same Adam step on every replica": rank r sums slice r of every rank's gradients (in the
    NVSwitch when multicast is available), updates slice r of every parameter tensor with ITS shard of
    exp_avg / exp_avg_sq and writes the new parameters into all replicas.

    params / grads: FlatGradBuffer objects whose flat tensors live in symmetric memory (`SymmetricFlat`; the gradient
    buffer of `ViewStreams(peer_group=...)` does), with a field of the same name and size for every group.
    groups: {field name: {'lr': ..., optional 'lr_hi', 'period', 'split'}} as for FusedAdam.  Arithmetic =
    torch.optim.Adam (the same device function as dmgs_adam_step).  Gradients are not cleared (ViewStreams.begin()
    does that)."""

    MULTICAST_MIN_WORLD = PeerAllReduce.MULTICAST_MIN_WORLD

    def __init__(self, params: FlatGradBuffer, param_sym: SymmetricFlat, grads: FlatGradBuffer, grad_sym, groups: Dict[str, dict],
                 betas=(0.9, 0.999), eps: float = 1e-8):
        import ctypes as C
        from . import _lib as L
        self.params, self.grads, self.psym, self.gsym = params, grads, param_sym, grad_sym
        self.world, self.rank = int(param_sym.world), int(param_sym.rank)
        if (int(grad_sym.world), int(grad_sym.rank)) != (self.world, self.rank):
            raise ValueError("parameter and gradient buffers belong to different groups")
        self.betas, self.eps, self.step_count = tuple(betas), float(eps), 0
        self.groups = {k: dict(v) for k, v in groups.items()}
        self.state: Dict[str, dict] = {}
        dev = params.flat.device
        for name in self.groups:
            if name not in params.offsets or name not in grads.offsets:
                raise KeyError(f"group {name!r} is not a field of both buffers")
            (po, pn), (go, gn) = params.offsets[name], grads.offsets[name]
            if pn != gn:
                raise ValueError(f"field {name!r}: {pn} parameters but {gn} gradients")
            b, e = C.c_int64(), C.c_int64()
            L.check(L.lib().dmgs_adam_exchange_shard(pn, self.world, self.rank, C.byref(b), C.byref(e)), "dmgs_adam_exchange_shard")
            n = 4 * (e.value - b.value)
            self.state[name] = {"begin": 4 * b.value, "end": 4 * e.value,
                                "exp_avg": torch.zeros(max(n, 4), dtype=torch.float32, device=dev),
                                "exp_avg_sq": torch.zeros(max(n, 4), dtype=torch.float32, device=dev)}

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, use_multicast: Optional[bool] = None):
        import ctypes as C
        from . import _lib as L
        self.step_count += 1
        names = list(self.groups)
        segs = (L.AdamXSegment * len(names))()
        for i, name in enumerate(names):
            g, st = self.groups[name], self.state[name]
            segs[i] = L.AdamXSegment(self.grads.offsets[name][0], self.params.offsets[name][0], self.params.offsets[name][1],
                                     st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), float(g["lr"]),
                                     float(g.get("lr_hi", g["lr"])), int(g.get("period", 0)), int(g.get("split", 0)))
        if use_multicast is None:
            use_multicast = self.world >= self.MULTICAST_MIN_WORLD
        mc = bool(use_multicast and self.psym.multicast_ptr and self.gsym.multicast_ptr)
        dev = self.params.flat.device
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        self.gsym.hdl.barrier(channel=0)  # every rank's gradients are complete, nobody still reads the old parameters
        for k in range(0, len(names), 8):  # DMGS_ADAM_MAX_SEGMENTS tensors per launch
            part = (L.AdamXSegment * len(names[k:k + 8]))(*segs[k:k + 8])
            L.check(L.lib().dmgs_adam_exchange_peer(self.world, self.rank, len(part), part, self.gsym.ptrs,
                                                    C.c_void_p(self.gsym.multicast_ptr) if mc else None, self.psym.ptrs,
                                                    C.c_void_p(self.psym.multicast_ptr) if mc else None, self.betas[0], self.betas[1],
                                                    self.eps, self.step_count, float(grad_scale), stream), "dmgs_adam_exchange_peer")
        self.gsym.hdl.barrier(channel=1)  # every slice of the new parameters has reached every replica
        return self.params

    def state_dict(self) -> dict:
        """This rank's SHARD of the state (ranges in elements of each field)."""
        return {"step": self.step_count, "betas": self.betas, "eps": self.eps, "world": self.world, "rank": self.rank,
                "groups": {k: dict(v) for k, v in self.groups.items()},
                "state": {k: {"begin": v["begin"], "end": v["end"], "exp_avg": v["exp_avg"].clone(),
                              "exp_avg_sq": v["exp_avg_sq"].clone()} for k, v in self.state.items()}}

    def load_state_dict(self, sd: dict):
        if (sd["world"], sd["rank"]) != (self.world, self.rank):
            raise ValueError("state shard of another rank / world size")
        self.step_count = int(sd["step"])
        for k, v in sd["state"].items():
            self.state[k]["exp_avg"].copy_(v["exp_avg"])
            self.state[k]["exp_avg_sq"].copy_(v["exp_avg_sq"])


def shard_rows(n_rows: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Row range [lo, hi) of an [n_rows, ...] tensor that `rank` stages, and the per-rank chunk (rows, the same on
    every rank: all-gather needs equal pieces, so the last chunks may be short or empty)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    chunk = (n_rows + world - 1) // world
    lo = min(rank * chunk, n_rows)
    return lo, min(lo + chunk, n_rows), chunk


class StagedInputs:
    """Double-buffered host -> device staging of a step's inputs on a side stream.

    `host` holds PINNED tensors.  `prefetch(slot)` enqueues the copy of all tensors into device buffer
    set `slot` on the copy stream; `acquire(slot)` makes the compute stream wait for that copy and
    returns the device tensors; `release(slot)` marks the buffers free once the compute stream has
    consumed them.  With two slots the copy of step i+1 overlaps the kernels of step i (PCIe and the
    SMs are independent engines), so a training loop whose inputs arrive from the host every step is
    not serialised behind the H2D transfer.

    group (ranks of ONE box holding identical host tensors -- the replicated Gaussians of the view-partitioned
    step): every rank copies only its 1/N row range of every tensor over PCIe and the pieces are all-gathered
    in place over NVLink (`dist.all_gather_into_tensor` on the copy stream), so the box moves the inputs
    over its host links once per step instead of once per rank."""

    def __init__(self, host: Dict[str, torch.Tensor], device, slots: int = 2, group=None):
        self.host = host
        self.group = group
        self.world = dist.get_world_size(group) if group is not None else 1
        self.rank = dist.get_rank(group) if group is not None else 0
        self.dev, self._full, self.plan = [], [], {}
        for k, v in host.items():
            self.plan[k] = shard_rows(v.shape[0], self.world, self.rank)
        for _ in range(slots):
            full = {k: torch.empty((self.plan[k][2] * self.world,) + tuple(v.shape[1:]), dtype=v.dtype, device=device)
                    for k, v in host.items()}
            self._full.append(full)  # padded to world equal chunks; the tensors handed out are the first n_rows
            self.dev.append({k: full[k][:v.shape[0]] for k, v in host.items()})
        self.copy_stream = torch.cuda.Stream(device=device)
        self.ready: List[Optional[torch.cuda.Event]] = [None] * slots
        self.freed: List[Optional[torch.cuda.Event]] = [None] * slots
        # bytes this rank copies host -> device per step
        self.bytes_per_step = sum((self.plan[k][1] - self.plan[k][0]) * (v[0].numel() if v.shape[0] else 0) * v.element_size()
                                  for k, v in host.items())

    def prefetch(self, slot: int):
        with torch.cuda.stream(self.copy_stream):
            if self.freed[slot] is not None:
                self.copy_stream.wait_event(self.freed[slot])
            for k, v in self.host.items():
                lo, hi, chunk = self.plan[k]
                if hi > lo:
                    self._full[slot][k][lo:hi].copy_(v[lo:hi], non_blocking=True)
            if self.world > 1:
                for k in self.host:
                    chunk, full = self.plan[k][2], self._full[slot][k]
                    dist.all_gather_into_tensor(full, full[self.rank * chunk:(self.rank + 1) * chunk], group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            self.ready[slot] = ev

    def acquire(self, slot: int) -> Dict[str, torch.Tensor]:
        if self.ready[slot] is None:
            self.prefetch(slot)
        torch.cuda.current_stream().wait_event(self.ready[slot])
        return self.dev[slot]

    def release(self, slot: int):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.freed[slot] = ev
        self.ready[slot] = None
