#!/usr/bin/env python
"""Kernel timeline of the drop-in module path with host inputs (bench.py's e2e.value step): 8 views one after the other
through GaussianRasterizer, autograd accumulating .grad, inputs staged from pinned memory, loss read back.
    python scripts/trace_module.py OUT_DIR"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dmgs_b200
from dmgs_b200 import GaussianRasterizationSettings, GaussianRasterizer, multiview as MV, synthetic as S

out = sys.argv[1]
os.makedirs(out, exist_ok=True)
dev = torch.device("cuda", 0)
dmgs_b200.configure(async_binning=True)
P, W, H, V = 1_000_000, 800, 800, 8
cl = S.random_cloud(P, seed=0, extent=1.3, log_scale_mean=math.log(0.01))
names = ["means3D", "scales", "rotations", "opacities", "shs"]
host = {k: cl[k].pin_memory() for k in names}
cams = [S.nerf_synthetic_camera(v, W, H).to(dev) for v in range(V)]
bg = torch.zeros(3, device=dev)
settings = [GaussianRasterizationSettings(H, W, math.tan(c.FoVx / 2), math.tan(c.FoVy / 2), bg, 1.0, c.world_view_transform,
                                          c.full_proj_transform, 3, c.camera_center, False, False) for c in cams]
gen = torch.Generator().manual_seed(77)
dLs = [torch.randn(3, H, W, generator=gen).to(dev) for _ in range(V)]
staged = MV.StagedInputs(host, dev)


def step(i, last):
    slot = i & 1
    bufs = staged.acquire(slot)
    if not last:
        staged.prefetch(slot ^ 1)
    t = {k: bufs[k].detach().requires_grad_() for k in names}
    loss = torch.zeros((), device=dev)
    for j in range(V):
        ras = GaussianRasterizer(settings[j])
        m2d = torch.zeros_like(t["means3D"], requires_grad=True)
        img, radii = ras(means3D=t["means3D"], means2D=m2d, shs=t["shs"], colors_precomp=None, opacities=t["opacities"],
                         scales=t["scales"], rotations=t["rotations"], cov3D_precomp=None)
        l = (img * dLs[j]).sum()
        l.backward()
        loss = loss + l.detach()
    h = float(loss.cpu())
    dmgs_b200.check_async()
    staged.release(slot)
    return h


for i in range(3):
    step(i, False)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(2):
        step(i + 1, i == 1)
    torch.cuda.synchronize()
trace = os.path.join(out, "trace_module_full.json")
prof.export_chrome_trace(trace)
ev = json.load(open(trace))["traceEvents"]
os.remove(trace)
ks = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e], key=lambda e: e["ts"])
t0 = ks[0]["ts"]
rows = [{"name": e["name"][:70], "stream": e["args"].get("stream"), "ts": round(e["ts"] - t0, 2), "dur": round(e["dur"], 2)} for e in ks]
json.dump(rows, open(os.path.join(out, "kernels_module.json"), "w"))
tot = {}
for r in rows:
    k = r["name"].split("(")[0].replace("void ", "").replace("dmgs::", "").replace("at::native::", "")[:48]
    a = tot.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += r["dur"]
end = max(r["ts"] + r["dur"] for r in rows)
print("span_us", round(end, 1), "per frame", round(end / 16, 1))
for k, (n, t) in sorted(tot.items(), key=lambda x: -x[1][1])[:22]:
    print(f"{t:9.1f} us {n:4d} x {t / n:7.1f}  {k}")
