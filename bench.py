#!/usr/bin/env python
"""bench.py -- fwd+bwd frames/s of the splatting hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload h0|c1|c3|c4]

Workload (N=1 default) = H0, the shape the metric is quoted on: 1 M random Gaussians, SH degree 3,
`shs`+`scales`+`rotations` mode, 800x800 NeRF-synthetic cameras.  A step renders
`views_per_rank` (8) views forward+backward per rank against seeded dL/dimage tensors and sums
their per-Gaussian gradients into one flat buffer; with N > 1 ranks the views of a step are
partitioned over the ranks (weak scaling: 8 views per rank) and the flat gradient buffer is
all-reduced over NCCL once per step.  `value` = frames (views) per second over all ranks.

Reference arm (`--impl reference`): the reference's own rasteriser is an un-vendored submodule
that is neither in /root/reference nor on the GPU box, so this arm times the CPU oracle (a port of
the same splat math, oracle/splat_oracle.c, all host threads) on one frame per step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fwd+bwd frames/sec @1M Gaussians 800x800"
WORKLOADS = {
    # name: (P, W, H, camera kind, extent, log-scale mean)
    "h0": (1_000_000, 800, 800, "nerf", 1.3, math.log(0.01)),
    "c1": (100_000, 800, 800, "nerf", 1.3, math.log(0.01)),
    "c3": (1_000_000, 1245, 825, "bicycle", 3.0, math.log(0.008)),
    # BASELINE.json configs[3]: 3 M Gaussians garden-scale, a 64-view batch partitioned across the ranks
    # (strong scaling: the batch is fixed, 64 / N views per rank), ring of bicycle-shaped cameras
    "c4": (3_000_000, 1245, 825, "ring64", 5.0, math.log(0.008)),
}
C4_TOTAL_VIEWS = 64
# Per-launch hardware counters of the kernels that can dominate (instructions executed, shared-memory wavefronts,
# DRAM bytes, issue-active %), read from the committed table the ncu --set full capture of THIS build was reduced
# to (scripts/ncu_counters.py -> profiles/r2_kernel_counters.json).  They are properties of the workload (same
# seed, same view), so dividing them by the LIVE CUDA-event duration gives live rates.
COUNTERS_FILE = os.path.join(ROOT, "profiles", "r2_kernel_counters.json")
VIEWS_PER_RANK = int(os.environ.get("DMGS_BENCH_VIEWS", "8"))
N_STREAMS = int(os.environ.get("DMGS_BENCH_STREAMS", "4"))


def kernel_counters(workload):
    try:
        with open(COUNTERS_FILE) as fh:
            tab = json.load(fh)
        return tab.get("workloads", {}).get(workload, {}), tab.get("source")
    except (OSError, ValueError):
        return {}, None


_RING = {}


def make_camera(kind, idx, W, H):
    from dmgs_b200 import synthetic as S
    if kind == "ring64":
        if (W, H) not in _RING:
            _RING[(W, H)] = S.ring_cameras(C4_TOTAL_VIEWS, W, H, radius=6.0, fovx=0.9, seed=0)
        return _RING[(W, H)][idx % C4_TOTAL_VIEWS]
    return S.nerf_synthetic_camera(idx, W, H) if kind == "nerf" else S.bicycle_camera(idx, W, H)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled every ~10 ms (NVML, in-process thread) while the timed region
    runs; falls back to `nvidia-smi -lms` when NVML cannot be loaded."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake"}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.h, self.stop_flag, self.thread = index, [], None, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((mhz, rs))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read_smi, daemon=True).start()
        except OSError:
            self.proc = None

    def _read_smi(self):
        for ln in self.proc.stdout:
            c = [x.strip() for x in ln.split(",")]
            if c and c[0].replace(".", "").isdigit():
                bits = sum(b for b, col in ((0x8, 2), (0x40, 3), (0x20, 4), (0x4, 5)) if len(c) > col and c[col] == "Active")
                self.max_mhz = float(c[1]) if len(c) > 1 and c[1].replace(".", "").isdigit() else None
                self.rows.append((float(c[0]), bits))

    def stop(self):
        if self.h is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for _, b in self.rows:
            bits |= b
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": getattr(self, "max_mhz", None),
                "reasons": [n for b, n in self.REASONS.items() if bits & b], "samples": len(sm),
                "source": "nvml, 10 ms period" if self.h is not None else "nvidia-smi -lms 100"}


def oracle_frame(O, pr, cl_np, dL):
    fwd = O.render_forward(pr, cl_np["means3D"], cl_np["opacities"], scales=cl_np["scales"],
                           rotations=cl_np["rotations"], shs=cl_np["shs"])
    O.render_backward(pr, fwd, dL, cl_np["means3D"], scales=cl_np["scales"], rotations=cl_np["rotations"],
                      shs=cl_np["shs"])
    return fwd


def cpu_baseline(workload, frames, warm=1):
    """Times the CPU oracle (port of the same splat math) on `frames` frames of the workload."""
    import torch
    from dmgs_b200 import synthetic as S
    from oracle import oracle as O
    P, W, H, kind, extent, lsm = WORKLOADS[workload]
    flags = O.use_native()  # the same source rebuilt with -O3 -march=native for the cores it is timed on
    # every core this process may use, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
    try:
        O.set_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        O.set_threads(os.cpu_count() or 1)
    cl = S.random_cloud(P, seed=0, extent=extent, log_scale_mean=lsm)
    cl_np = {k: v.numpy() for k, v in cl.items()}
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(77)).numpy()
    times = []
    for f in range(warm + frames):
        cam = make_camera(kind, f, W, H)
        pr = O.make_params(P, W, H, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2), [0, 0, 0],
                           cam.world_view_transform.numpy(), cam.full_proj_transform.numpy(),
                           cam.camera_center.numpy())
        t0 = time.perf_counter()
        oracle_frame(O, pr, cl_np, dL)
        times.append(time.perf_counter() - t0)
    t = sum(times[warm:])
    return {"value": frames / t, "unit": "frames/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"{frames} full frames (fwd+bwd) of workload {workload} ({P} Gaussians, {W}x{H}), "
                      f"OpenMP over Gaussians/tiles, {warm} warm-up frame; gcc {flags}"}, t / frames


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, Wm = max(args.steps, 1), max(args.warmup, 0)
    P, W, H, kind, _, _ = WORKLOADS[args.workload]
    cb, sec_per_frame = cpu_baseline(args.workload, K, warm=min(Wm, 2))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus,
            "steps": K, "warmup": Wm, "ms_per_step": 1e3 * sec_per_frame, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload.upper()}: {P} random Gaussians SH-3, {W}x{H}, shs+scales+rotations, "
                                   "fwd+bwd; reference arm = CPU oracle (reference rasteriser sources/install unavailable), "
                                   "one frame per step"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import dmgs_b200
    from dmgs_b200 import GaussianRasterizationSettings, GaussianRasterizer, _lib as L, multiview as MV, synthetic as S
    from dmgs_b200.rasterizer import rasterize_backward, rasterize_forward
    # sync-free binning: no host read-back of the instance count inside a frame; the overflow flags are
    # checked once per step, next to the step's other host synchronisation (a failed step is repeated)
    dmgs_b200.configure(async_binning=not args.sync_binning)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU (there is no CPU path in the product)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()

    P, W, H, kind, extent, lsm = WORKLOADS[args.workload]
    K, Wm = args.steps, max(args.warmup, 3)
    strong = args.workload == "c4"  # a fixed 64-view batch split over the ranks; everything else: 8 views per rank
    views_per_rank = (C4_TOTAL_VIEWS // world) if strong else VIEWS_PER_RANK
    cl = S.random_cloud(P, seed=0, extent=extent, log_scale_mean=lsm)  # replicated on every rank
    names = ["means3D", "scales", "rotations", "opacities", "shs"]
    host = {k: cl[k].pin_memory() for k in names}
    d = {k: host[k].to(dev) for k in names}
    n_views = views_per_rank * world
    cams = [make_camera(kind, v, W, H).to(dev) for v in range(n_views)]
    my_views = MV.partition_views(n_views, world, rank)
    bg = torch.zeros(3, device=dev)
    settings = [GaussianRasterizationSettings(H, W, math.tan(c.FoVx / 2), math.tan(c.FoVy / 2), bg, 1.0,
                                              c.world_view_transform, c.full_proj_transform, 3, c.camera_center,
                                              False, False) for c in cams]
    gen = torch.Generator().manual_seed(77)
    dLs = [torch.randn(3, H, W, generator=gen).to(dev) for _ in range(min(n_views, 8))]
    # flat per-Gaussian gradient buffer (62 floats per Gaussian): the all-reduce payload
    # views of a step run round-robin on N_STREAMS CUDA streams (MV.ViewStreams), one accumulator each
    # N > 1: the summed buffer lives in symmetric memory and is exchanged over NVLink peer memory
    # (dmgs_allreduce_peer; DMGS_BENCH_ALLREDUCE=nccl forces torch.distributed.all_reduce)
    use_peer = world > 1 and os.environ.get("DMGS_BENCH_ALLREDUCE", "peer") != "nccl"
    deferred = os.environ.get("DMGS_BENCH_DEFERRED_SH", "1") == "1"
    vs = MV.ViewStreams(P, MV.RASTER_WIDTHS_SH, dev, n=N_STREAMS, peer_group=dist.group.WORLD if use_peer else None,
                        deferred_sh_views=views_per_rank if deferred else 0)
    if world == 1:
        allreduce_kind = "none (single GPU)"
    elif vs.peer is not None:
        mm = vs.peer.multicast_ptr and world >= vs.peer.MULTICAST_MIN_WORLD
        allreduce_kind = "NVLink peer memory, " + ("NVSwitch multimem.ld_reduce/st" if mm else "P2P loads/stores")
    else:
        allreduce_kind = "NCCL all-reduce" + (f" (peer memory unavailable: {vs.peer_error})" if vs.peer_error else "")
    stage_ms, ev_log = {}, []

    def hook_factory(events):
        def hook(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events.append((name, e))
        return hook

    stats = {"R": 0, "frames": 0, "redone": 0, "host_s": 0.0}
    # views as CUDA graphs (multiview.ViewStreams.capture): captured during the warm-up steps, after one eager step
    GRAPHS = os.environ.get("DMGS_BENCH_GRAPHS", "1") == "1"
    graphs = {"ready": False, "captures": 0}
    r_dev = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(N_STREAMS)]

    def one_view(j, v, record, acc):
        rec = vs.sh_record(j, settings[v].campos)
        events = []
        if record:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
            events.append(("start", e0))
        hook = hook_factory(events) if record else None
        color, radii, st = rasterize_forward(settings[v], d["means3D"], d["opacities"], d["shs"], None,
                                             d["scales"], d["rotations"], None, stage_hook=hook)
        rasterize_backward(st, dLs[j % len(dLs)], d["means3D"], d["shs"], d["scales"], d["rotations"], None, False,
                           stage_hook=hook, accumulate_into=acc, sh_record=rec, verify=False)
        if st._count is None:  # sync-free frame: the count is still on the device (same stream: ordered)
            r_dev[j % N_STREAMS].add_(st.ws.meta[0:1])
        else:
            stats["R"] += st.num_rendered
        if record:
            ev_log.append(events)

    def step(record, single=False, exchange=True):
        # single=True: every view on the current stream into the first accumulator (the stage-timing steps
        # after the timed region; record=True puts CUDA events between the stages)
        th = time.perf_counter()
        if single:
            vs.begin()
            for j, v in enumerate(my_views):
                one_view(j, v, record, vs.buf.views)
            if deferred:
                vs.finish(d["means3D"], d["shs"], 3)
        else:
            vs.begin()
            for j, v in enumerate(my_views):
                if graphs["ready"] and not vs.captured(j):
                    try:
                        vs.capture(j, lambda acc, j=j, v=v: one_view(j, v, False, acc))
                        graphs["captures"] += 1
                    except Exception as e:  # capture unavailable: the eager path is the same work
                        graphs["ready"], graphs["error"] = False, f"{type(e).__name__}: {e}"[:200]
                        vs.drop_graphs()
                if graphs["ready"]:  # the whole view (forward + backward, ~25 launches) is ONE graph launch
                    vs.replay(j)
                else:
                    vs.run(j, lambda acc, j=j, v=v: one_view(j, v, False, acc))
            vs.finish(d["means3D"], d["shs"], 3)
        stats["frames"] += len(my_views)
        if world > 1 and exchange:
            vs.all_reduce_()
        stats["host_s"] += time.perf_counter() - th  # host time to ENQUEUE the step (no synchronisation so far)
        ok = dmgs_b200.check_async()
        ok = vs.poll_captured() and ok
        if not single and GRAPHS and not args.sync_binning and "error" not in graphs:
            graphs["ready"] = True  # capacities and host-side camera values exist after the first eager step
        if not ok:  # a frame overflowed its binning buffer: the step does not count
            stats["redone"] += 1
            if record:
                del ev_log[-len(my_views):]
            stats["frames"] -= len(my_views)
            step(record, single, exchange)

    for _ in range(Wm):
        step(False)
    torch.cuda.synchronize()
    stats.update(R=0, frames=0, redone=0, host_s=0.0)
    for t in r_dev:
        t.zero_()
    launches0 = lib.dmgs_launch_count() + vs.replayed_kernel_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(K):
        step(False)
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = t0.elapsed_time(t1)
    launches = lib.dmgs_launch_count() + vs.replayed_kernel_launches - launches0  # eager launches + graph kernel nodes
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tm = torch.tensor([ms], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
        lt = torch.tensor([float(launches)], device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    frames_timed, redone_timed = stats["frames"], stats["redone"]
    host_enqueue_ms = 1e3 * stats["host_s"] / max(K, 1)

    # ---- multi-GPU correctness, outside the timed region (driver-visible: keys of the JSON line)
    mgpu = None
    if world > 1:
        mgpu = multi_gpu_checks(torch, dist, MV, vs, dev, world, rank, P, n_views, settings, d, dLs, step,
                                rasterize_forward, rasterize_backward)
        if vs.peer is not None and not args.quick:
            mgpu["optimizer_tail"] = optimizer_tail_timing(torch, dist, MV, vs, dev, world, P, d, names)

    # per-stage kernel durations: two more steps on ONE stream with CUDA events between the stages (with
    # several views in flight the stages of different views overlap and cannot be timed individually)
    step(False, single=True)  # the caching allocator's per-stream pools: first single-stream step allocates
    torch.cuda.synchronize()
    for _ in range(2):
        step(True, single=True)
    torch.cuda.synchronize()
    for events in ev_log:
        for (n0, a), (n1, b) in zip(events[:-1], events[1:]):
            stage_ms[n1] = stage_ms.get(n1, 0.0) + a.elapsed_time(b)
    for k in stage_ms:
        stage_ms[k] /= max(len(ev_log), 1)
    Ravg = (stats["R"] + sum(int(t.item()) for t in r_dev)) / max(stats["frames"] + len(my_views) * stats["redone"], 1)
    stats["redone"] = redone_timed
    value = (views_per_rank * world * K) / (ms / 1e3)

    # ---- the DROP-IN path, device resident: what the reference's trainers would see (one view per iteration
    # through the GaussianRasterizer nn.Module, autograd backward, zero_grad(set_to_none=True);
    # train_geo_stage2.py:91-132 / gaussian_renderer/__init__.py:18-101).  Timed with CUDA events over >= 2 s.
    def dropin_loop(n_frames):
        t = {k: d[k].detach().requires_grad_() for k in names}
        for i in range(n_frames):
            v = my_views[i % len(my_views)]
            ras = GaussianRasterizer(settings[v])
            m2d = torch.zeros_like(t["means3D"], requires_grad=True)  # screenspace_points of render()
            img, radii = ras(means3D=t["means3D"], means2D=m2d, shs=t["shs"], colors_precomp=None,
                             opacities=t["opacities"], scales=t["scales"], rotations=t["rotations"], cov3D_precomp=None)
            img.backward(dLs[i % len(dLs)])
            for x in t.values():
                x.grad = None

    def time_dropin(sync_free):
        dmgs_b200.configure(async_binning=sync_free)
        try:
            dropin_loop(2 * len(my_views))
            torch.cuda.synchronize()
            n = max(views_per_rank * K, 64)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dropin_loop(n)
            b.record()
            torch.cuda.synchronize()
            dmgs_b200.check_async()
            return n / (a.elapsed_time(b) / 1e3)
        finally:
            dmgs_b200.configure(async_binning=not args.sync_binning)

    full = not args.quick  # --quick (kernel-variant experiments): `value` and the stage times only
    value_dropin = time_dropin(False) if world == 1 and full else None          # default settings: upstream's host read-back
    value_dropin_syncfree = time_dropin(True) if world == 1 and full else None  # configure(async_binning=True)

    # ---- end-to-end with HOST inputs (rank-local; max over ranks).  Every step copies all inputs from
    # pinned host memory (double-buffered on a copy stream, so the copy of step i+1 overlaps the kernels
    # of step i) and reads the step's loss back to the host.  Two public entry points are measured:
    #   "autograd_module" (= e2e.value): the drop-in GaussianRasterizer nn.Module, one view after the other,
    #                    autograd accumulating .grad over the views of the step
    #   "training_step": dmgs_b200.multiview (ViewStreams + accumulate_view over the C ABI) -- the same call
    #                    sequence as the `value` region, plus the copies and the loss
    # N > 1: the inputs are replicated, so every rank copies 1/N of every tensor over PCIe and the pieces are
    # all-gathered over NVLink on the copy stream (its own communicator: never queued behind the gradient exchange)
    stage_group = dist.new_group() if world > 1 else None
    staged = MV.StagedInputs(host, dev, group=stage_group)

    def e2e_step_module(i, last):
        slot = i & 1
        bufs = staged.acquire(slot)
        if not last:
            staged.prefetch(slot ^ 1)
        t = {k: bufs[k].detach().requires_grad_() for k in names}
        loss = torch.zeros((), device=dev)
        try:
            for j, v in enumerate(my_views):
                ras = GaussianRasterizer(settings[v])
                m2d = torch.zeros_like(t["means3D"], requires_grad=True)
                img, radii = ras(means3D=t["means3D"], means2D=m2d, shs=t["shs"], colors_precomp=None,
                                 opacities=t["opacities"], scales=t["scales"], rotations=t["rotations"],
                                 cov3D_precomp=None)
                l = (img * dLs[j % len(dLs)]).sum()
                l.backward()
                loss = loss + l.detach()
        except dmgs_b200.rasterizer.BinningOverflowError:  # capacity raised: repeat the step on the same inputs
            staged.ready[slot] = torch.cuda.Event()
            staged.ready[slot].record()
            return e2e_step_module(i, last)
        if world > 1:
            g = torch.cat([t[k].grad.reshape(-1) for k in names])
            dist.all_reduce(g)
        host_loss = float(loss.cpu())  # device -> host read of the step's result
        dmgs_b200.check_async()
        staged.release(slot)
        return host_loss

    def e2e_step_training(i, last):
        slot = i & 1
        bufs = staged.acquire(slot)
        if not last:
            staged.prefetch(slot ^ 1)
        inputs = {k: bufs[k] for k in names}
        vs.begin()
        losses = []

        def view(j, v, acc):  # forward, loss, backward of one view on the staged inputs of this slot
            dl = dLs[j % len(dLs)]
            rec = vs.sh_record(j, settings[v].campos)
            return MV.accumulate_view(settings[v], inputs, lambda img: ((img * dl).sum(), dl), acc, sh_record=rec)[0]

        for j, v in enumerate(my_views):
            key = 1000 * (slot + 1) + j  # one graph per (staging slot, view): the slot's device buffers keep their addresses
            if graphs["ready"] and not vs.captured(key):
                try:
                    vs.capture(key, lambda acc, j=j, v=v: view(j, v, acc))
                    graphs["captures"] += 1
                except Exception as e:
                    graphs["ready"], graphs["error"] = False, f"{type(e).__name__}: {e}"[:200]
                    vs.drop_graphs()
            if graphs["ready"]:
                losses.append(vs.replay(key))  # the captured callable's result: the view's loss (a static 0-d tensor)
            else:
                losses.append(vs.run(j, lambda acc, j=j, v=v: view(j, v, acc)))
        vs.finish(inputs["means3D"], inputs["shs"], 3)
        if world > 1:
            vs.all_reduce_()
        host_loss = float(torch.stack(losses).sum().cpu())  # device -> host read of the step's result
        ok = dmgs_b200.check_async()
        if not (vs.poll_captured() and ok):
            staged.ready[slot] = torch.cuda.Event()
            staged.ready[slot].record()
            return e2e_step_training(i, last)
        staged.release(slot)
        return host_loss

    Ke = max(3, K)  # the first step's copy is exposed (pipeline fill), later copies hide behind the previous step

    def time_e2e(step_fn):
        for i in range(2):
            step_fn(i, last=(i == 1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        te = time.perf_counter()
        for i in range(Ke):  # exactly Ke host->device copies of the full input set inside the timed region
            step_fn(i, last=(i == Ke - 1))
        torch.cuda.synchronize()
        dt = time.perf_counter() - te
        if world > 1:
            tm = torch.tensor([dt], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dt = float(tm.item())
        return (views_per_rank * world * Ke) / dt

    # gradient accumulation over the step's views inside the kernels (opt-in: see dmgs_b200.configure)
    FUSED_ACC = os.environ.get("DMGS_BENCH_FUSED_ACC", "1") == "1"
    dmgs_b200.configure(accumulate_grad_in_place=FUSED_ACC)
    e2e_module = time_e2e(e2e_step_module) if full else None
    dmgs_b200.configure(accumulate_grad_in_place=False)
    vs.drop_graphs()  # the graphs of the `value` region own one workspace each: free them before capturing the e2e ones
    e2e_training = time_e2e(e2e_step_training) if full else None
    vs.drop_graphs()
    h2d = staged.bytes_per_step
    # what the staging delivered must be the resident copy, bit for bit (sharded copy + all-gather at N > 1)
    torch.cuda.synchronize()
    staged_ok = all(bool(torch.equal(staged.dev[s_][k], d[k])) for s_ in range(2) for k in names) if full else None
    if world > 1 and full:
        ok_t = torch.tensor([1.0 if staged_ok else 0.0], device=dev)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        staged_ok = bool(ok_t.item() > 0.5)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines.  Single-kernel stages are timed exactly by the hooks.  The blend kernels are bound by the
    # SMs' instruction issue rate (DRAM 1-2 %), so their roof is warp-instructions/s; the per-Gaussian kernels
    # are HBM bound.  `roofline` describes the dominant kernel on ITS roof; `roofline_hbm` the dominant HBM kernel.
    peak, peak_src = peaks()
    T = ((W + 15) // 16) * ((H + 15) // 16)
    B_in, B_geo, B_gin = 236, 75, 232
    alg = {  # algorithmic bytes per frame for each stage
        "preprocess_sort_scan": P * (B_in + B_geo) + 4 * 16 * P + 8 * P,
        "binning": 8 * Ravg + 2 * 16 * Ravg + 4 * Ravg + 8 * T,
        "blend_fwd": 40 * Ravg + 20 * W * H,
        "blend_bwd": 40 * Ravg + 20 * W * H + 40 * P,
        # deferred SH gradient: no SH row is read or written by the per-view kernel (44 B of other inputs, the
        # state, the blend's 48-byte gradient record, 13 small gradients read-modify-written, a 16-byte record)
        "preprocess_bwd": P * (44 + B_geo + 48 + 2 * 52 + 16) if deferred else P * (B_in + B_geo + 40) + P * B_gin,
    }
    single = ["blend_fwd", "blend_bwd", "preprocess_bwd"]
    dom = max(single, key=lambda k: stage_ms.get(k, 0.0))
    stages = {k: {"ms": round(v, 4), "alg_GBps": round(alg[k] / (v * 1e-3) / 1e9, 1) if v > 0 else None}
              for k, v in stage_ms.items()}
    counters, counters_src = kernel_counters(args.workload)

    def hbm_roof(kname):
        sec = stage_ms.get(kname, 0.0) * 1e-3
        ach = alg[kname] / sec / 1e9 if sec > 0 else 0.0
        c = counters.get(kname, {})
        return {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": c.get("dram_bytes"), "peak_source": peak_src,
                "traffic_source": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, from capture {counters_src}"
                if c.get("dram_bytes") is not None else None}

    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
    issue_peak = n_sms * 4 * sm_mhz * 1e6 / 1e9  # G warp-instructions/s: one per scheduler and clock
    if dom in ("blend_fwd", "blend_bwd"):
        c = counters.get(dom, {})
        sec = stage_ms[dom] * 1e-3
        inst = c.get("inst_executed")
        ach = inst / sec / 1e9 if inst else None
        roofline = {"kernel": dom, "bound": "fp32_issue", "achieved": ach, "peak": issue_peak, "unit": "Gwarp-inst/s",
                    "frac": (ach / issue_peak) if ach else None,
                    "peak_source": f"{n_sms} SMs x 4 schedulers x {sm_mhz:.0f} MHz (clock sampled during the run)",
                    "inst_executed_per_launch": inst, "issue_active_pct_in_capture": c.get("issue_active_pct"),
                    "smem_wavefronts_per_s": (c["smem_wavefronts"] / sec) if c.get("smem_wavefronts") else None,
                    "warp_entry_pairs_per_s": (c["warp_entry_pairs"] / sec) if c.get("warp_entry_pairs") else None,
                    "traffic": c.get("dram_bytes"),
                    "counters_source": f"per-launch counters from capture {counters_src}; duration live (CUDA events)"
                    if c else "no committed capture for this workload",
                    "hbm": {k: v for k, v in hbm_roof(dom).items() if k in ("achieved", "peak", "frac", "unit")},
                    "note": "blend kernels read ~10 % of their algorithmic bytes from DRAM (only the first part of every "
                            "tile list is consumed, records hit L2): the HBM fraction is reported for the contract, the "
                            "issue rate is the roof (DESIGN.md section 5)"}
    else:
        roofline = hbm_roof(dom)
    hbm_dom = max(["preprocess_bwd"], key=lambda k: stage_ms.get(k, 0.0))
    cb = None
    if world == 1 and not args.no_cpu_baseline and full:
        cb, _ = cpu_baseline(args.workload, 3)
    wl = {"h0": "H0", "c1": "C1", "c3": "C3", "c4": "C4"}[args.workload]
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{wl}: {P} random Gaussians SH-3, {W}x{H}, shs+scales+rotations, fwd+bwd",
                   "views_per_rank_per_step": views_per_rank, "parallelism": f"views x{world} (replicated Gaussians, "
                   "one all-reduce of the flat gradient buffer per step)" if world > 1 else "single GPU",
                   "all_reduce": allreduce_kind,
                   "avg_instances_R": Ravg, "view_streams": N_STREAMS,
                   "blend_residency_ctas_per_sm": vs.blend_residency, "place_smem_kb_per_sm": vs.place_smem_kb,
                   "cuda_graphs": ("one graph per view (forward + backward), captured in the warm-up, replayed in the "
                                   f"timed region; {graphs['captures']} captures") if graphs["ready"] else graphs.get("error", "off"),
                   "sh_gradient": "deferred: 16-byte records per view, rows formed once per step (dmgs_sh_grad_expand)" if deferred
                   else "read-modify-write of the [P,16,3] rows every view",
                   "stage_timing": "2 single-stream steps right after the timed region, CUDA events between stages",
                   "binning": "host read-back of the instance count every frame" if args.sync_binning else
                   "sync-free (capacity from earlier frames, overflow flags checked once per step)",
                   "steps_repeated_after_overflow": stats["redone"],
                   "l2": f"inputs ({h2d / 1e6:.0f} MB) + state exceed the 126 MB L2; no flush needed"},
        "stages": stages,
        "roofline": roofline,
        "roofline_hbm": hbm_roof(hbm_dom),
        "value_dropin": value_dropin,
        "value_dropin_syncfree": value_dropin_syncfree,
        "value_dropin_api": "GaussianRasterizer nn.Module, ONE view per iteration, autograd backward, grads set to None "
                            "(what train_geo_stage2.py:91-132 would see), inputs resident, CUDA events; value_dropin = "
                            "default settings (host read-back of the instance count, as upstream), _syncfree = "
                            "configure(async_binning=True)",
        "e2e": {"value": e2e_module, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "steps": Ke, "api": "drop-in GaussianRasterizer nn.Module (the reference-facing call), the step's views one "
                "after the other, loss.backward() accumulating .grad over the views ("
                + ("configure(async_binning=True, accumulate_grad_in_place=True): from the second view on the kernels add "
                   "into the existing .grad instead of autograd's AccumulateGrad pass" if FUSED_ACC else
                   "configure(async_binning=True); autograd's AccumulateGrad adds") +
                "); host inputs staged from pinned memory every step, loss read back every step",
                "training_step": e2e_training,
                "training_step_api": "dmgs_b200.multiview training step (ViewStreams + accumulate_view over the C ABI; one CUDA "
                                     "graph per (staging slot, view) when graphs are on), "
                                     "same copies and read-back",
                "staging": ("every rank copies 1/N of every (replicated) input tensor from pinned host memory, the pieces "
                            "are all-gathered over NVLink on the copy stream; h2d_bytes_per_step is per rank" if world > 1
                            else "all inputs copied from pinned host memory on a copy stream, double-buffered"),
                "h2d_bytes_per_step_all_ranks": int(sum(v.numel() * v.element_size() for v in host.values())) if world > 1 else h2d,
                "staged_equals_resident": staged_ok},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(host_enqueue_ms, 3), "clocks": clocks,
    }
    if mgpu is not None:
        line.update(mgpu)
    if cb is not None:
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def optimizer_tail_timing(torch, dist, MV, vs, dev, world, P, d, names, iters=10):
    """The tail of a training step, outside the timed region: (a) dmgs_allreduce_peer followed by FusedAdam on every
    replica, (b) the two fused (dmgs_adam_exchange_peer: reduce-scatter + sharded Adam + parameter all-gather in
    one kernel per rank).  ms per step, CUDA events, max over ranks; the two paths must agree on the parameters."""
    from dmgs_b200.optim import FusedAdam
    widths = {k: tuple(d[k].shape[1:]) for k in names}
    lrs = {"means3D": {"lr": 1.6e-4}, "opacities": {"lr": 5e-2}, "scales": {"lr": 5e-3}, "rotations": {"lr": 1e-3},
           "shs": {"lr": 2.5e-3, "lr_hi": 2.5e-3 / 20, "period": 48, "split": 3}}
    palloc, ph = MV.SymmetricFlat.allocator(dev, dist.group.WORLD)
    params = MV.FlatGradBuffer(P, widths, dev, allocate=palloc)
    for k in names:
        params.views[k].copy_(d[k])
    replica = {k: d[k].clone() for k in names}
    fused = MV.ShardedPeerAdam(params, ph[0], vs.buf, vs.peer, {k: lrs[k] for k in names}, eps=1e-15)
    plain = FusedAdam([{"params": [replica[k]], "name": k, **lrs[k]} for k in names], lr=0.0, eps=1e-15)
    gen = torch.Generator(device=dev).manual_seed(4321 + dist.get_rank())
    grad0 = torch.zeros_like(vs.buf.flat)
    for k, view in vs.buf.views.items():
        o, m = vs.buf.offsets[k]
        grad0[o:o + m] = torch.randn(m, generator=gen, device=dev) * 1e-3
    scale = 1.0 / (8 * world)

    def timed(fn):
        ts = []
        for _ in range(iters + 2):
            vs.buf.flat.copy_(grad0)
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts = sorted(ts[2:])
        t = torch.tensor([ts[len(ts) // 2]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def two_kernels():
        vs.all_reduce_(scale)
        plain.step(grads={k: vs.buf.views[k] for k in names})

    ms_plain = timed(two_kernels)
    ms_fused = timed(lambda: fused.step(grad_scale=scale))
    err = max(float((params.views[k] - replica[k]).abs().max() / replica[k].abs().max()) for k in names)
    e = torch.tensor([err], device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    return {"allreduce_then_adam_ms": round(ms_plain, 4), "fused_exchange_adam_ms": round(ms_fused, 4),
            "params_max_rel_diff_after_%d_steps" % (iters + 2): float(e.item()),
            "optimizer_state_bytes_per_rank": {"replicated": 8 * sum(int(d[k].numel()) for k in names),
                                               "sharded": 8 * sum(int(v["exp_avg"].numel()) for v in fused.state.values())},
            "note": "tail of a training step: dmgs_allreduce_peer + dmgs_adam_step on every replica vs ONE "
                    "dmgs_adam_exchange_peer per rank (slice-owner reduces, applies Adam on its state shard, broadcasts parameters)"}


def multi_gpu_checks(torch, dist, MV, vs, dev, world, rank, P, n_views, settings, d, dLs, step, rasterize_forward,
                     rasterize_backward):
    """Outside the timed region: (i) the peer-memory all-reduce against NCCL on the same data; (ii) the N-rank
    reduced gradient buffer against rank 0 rendering ALL views alone (row accumulation, no deferred SH, one
    stream: an independent path).  Returns keys for the JSON line."""
    out = {}
    flat = vs.buf.flat
    # (i) same random payload on both paths
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    payload = torch.randn(flat.numel(), generator=g, device=dev)
    ref = payload.clone()
    dist.all_reduce(ref)
    flat.copy_(payload)
    torch.cuda.synchronize()
    dist.barrier()
    vs.all_reduce_()
    torch.cuda.synchronize()
    err = float(((flat - ref).abs().max() / ref.abs().max()).item())
    digest = flat.view(torch.int32).long().sum().reshape(1)  # order-independent checksum of the bit patterns
    digests = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(digests, digest)
    out["allreduce_check"] = {"max_rel_err_vs_nccl": err, "same_bits_across_ranks": bool(all(int(x) == int(digests[0]) for x in digests)),
                              "payload_floats": int(flat.numel()),
                              "path": "dmgs_allreduce_peer" if vs.peer is not None else "nccl (peer memory unavailable)"}
    # (ii) one normal N-rank step, then rank 0 alone over all views
    step(False)
    torch.cuda.synchronize()
    reduced = {k: v.clone() for k, v in vs.buf.views.items()}
    worst, per_field = 0.0, {}
    if rank == 0:
        solo = MV.FlatGradBuffer(P, MV.RASTER_WIDTHS_SH, dev)
        for v in range(n_views):
            j = v // world  # the local index this view has on its owning rank: the same dL/dimage
            color, radii, st = rasterize_forward(settings[v], d["means3D"], d["opacities"], d["shs"], None, d["scales"],
                                                 d["rotations"], None)
            rasterize_backward(st, dLs[j % len(dLs)], d["means3D"], d["shs"], d["scales"], d["rotations"], None, False,
                               accumulate_into=solo.views)
        torch.cuda.synchronize()
        for k, v in solo.views.items():
            den = float(torch.linalg.norm(v.double()))
            e = float(torch.linalg.norm((reduced[k] - v).double())) / max(den, 1e-30)
            per_field[k] = e
            worst = max(worst, e)
        del solo
    w = torch.tensor([worst], device=dev)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    out["nrank_vs_1rank_rel_err"] = float(w.item())
    if rank == 0:
        out["nrank_vs_1rank_per_field"] = per_field
    dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("DMGS_BENCH_WORKLOAD", "h0"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="experiments: skip the drop-in, end-to-end and CPU legs "
                    "(the line then carries nulls there and is NOT a bench result)")
    ap.add_argument("--sync-binning", action="store_true",
                    help="read the instance count back every frame (upstream behaviour) instead of sync-free binning")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
