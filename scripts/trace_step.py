#!/usr/bin/env python
"""Kernel timeline of the view-batched training step (H0, bench.py's step) through torch.profiler (CUPTI; nsys is not
in the image): which kernels run concurrently, how long the device idles, how the stages of different views
interleave.  Writes a compact per-kernel table (name, stream, start us, duration us) and a summary.

    python scripts/trace_step.py OUT_DIR [streams] [graphs 0|1] [steps]
"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dmgs_b200  # noqa: E402
from dmgs_b200 import GaussianRasterizationSettings, multiview as MV, synthetic as S  # noqa: E402
from dmgs_b200.rasterizer import rasterize_backward, rasterize_forward  # noqa: E402


def main():
    out = sys.argv[1]
    n_streams = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    graphs = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    os.makedirs(out, exist_ok=True)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dmgs_b200.configure(async_binning=True)
    P, W, H, V = 1_000_000, 800, 800, 8
    cl = S.random_cloud(P, seed=0, extent=1.3, log_scale_mean=math.log(0.01))
    names = ["means3D", "scales", "rotations", "opacities", "shs"]
    d = {k: cl[k].to(dev) for k in names}
    cams = [S.nerf_synthetic_camera(v, W, H).to(dev) for v in range(V)]
    bg = torch.zeros(3, device=dev)
    settings = [GaussianRasterizationSettings(H, W, math.tan(c.FoVx / 2), math.tan(c.FoVy / 2), bg, 1.0,
                                              c.world_view_transform, c.full_proj_transform, 3, c.camera_center,
                                              False, False) for c in cams]
    gen = torch.Generator().manual_seed(77)
    dLs = [torch.randn(3, H, W, generator=gen).to(dev) for _ in range(V)]
    vs = MV.ViewStreams(P, MV.RASTER_WIDTHS_SH, dev, n=n_streams, deferred_sh_views=V)

    def one_view(j, acc):
        rec = vs.sh_record(j, settings[j].campos)
        color, radii, st = rasterize_forward(settings[j], d["means3D"], d["opacities"], d["shs"], None, d["scales"],
                                             d["rotations"], None)
        rasterize_backward(st, dLs[j], d["means3D"], d["shs"], d["scales"], d["rotations"], None, False,
                           accumulate_into=acc, sh_record=rec, verify=False)

    ready = [False]

    def step():
        vs.begin()
        for j in range(V):
            if graphs and ready[0]:  # one CUDA graph per view (ViewStreams.capture), as bench.py does
                if not vs.captured(j):
                    vs.capture(j, lambda acc, j=j: one_view(j, acc))
                vs.replay(j)
            else:
                vs.run(j, lambda acc, j=j: one_view(j, acc))
        vs.finish(d["means3D"], d["shs"], 3)
        ok = dmgs_b200.check_async()
        ok = vs.poll_captured() and ok
        ready[0] = True
        if not ok:
            step()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
    trace = os.path.join(out, "trace_full.json")
    prof.export_chrome_trace(trace)
    ev = json.load(open(trace))["traceEvents"]
    os.remove(trace)
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
    ks.sort(key=lambda e: e["ts"])
    t0 = ks[0]["ts"]
    rows = [{"name": e["name"][:60], "stream": e["args"].get("stream"), "ts": round(e["ts"] - t0, 2), "dur": round(e["dur"], 2)}
            for e in ks]
    json.dump(rows, open(os.path.join(out, f"kernels_s{n_streams}_g{int(graphs)}.json"), "w"))
    # summary: busy time (union of intervals), concurrency-weighted time, per-kernel totals
    end = max(r["ts"] + r["dur"] for r in rows)
    pts = sorted([(r["ts"], 1) for r in rows] + [(r["ts"] + r["dur"], -1) for r in rows])
    busy, depth, last, conc = 0.0, 0, 0.0, {}
    for t, dlt in pts:
        if depth > 0:
            busy += t - last
        conc[depth] = conc.get(depth, 0.0) + (t - last)
        depth += dlt
        last = t
    tot = {}
    for r in rows:
        k = r["name"].split("(")[0].replace("void ", "").replace("dmgs::", "")[:40]
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += r["dur"]
    summ = {"streams": n_streams, "cuda_graphs": graphs, "steps": steps, "span_us": round(end, 1),
            "us_per_frame": round(end / (steps * V), 1), "busy_us": round(busy, 1), "idle_us": round(end - busy, 1),
            "time_at_concurrency_us": {str(k): round(v, 1) for k, v in sorted(conc.items())},
            "kernel_totals_us": {k: [n, round(t, 1), round(t / n, 1)] for k, (n, t) in sorted(tot.items(), key=lambda x: -x[1][1])}}
    json.dump(summ, open(os.path.join(out, f"summary_s{n_streams}_g{int(graphs)}.json"), "w"), indent=1)
    print(json.dumps(summ)[:3000])


if __name__ == "__main__":
    main()
