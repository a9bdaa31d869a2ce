"""Breakdown of the end-to-end step of bench.py (H0): pinned H2D alone, the autograd module path with
resident inputs, and both together.  python scripts/e2e_probe.py"""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dmgs_b200
from dmgs_b200 import GaussianRasterizationSettings, GaussianRasterizer, multiview as MV, synthetic as S

dmgs_b200.configure(async_binning=True)
dev = torch.device("cuda", 0)
P, W, H = 1_000_000, 800, 800
cl = S.random_cloud(P, seed=0, extent=1.3, log_scale_mean=math.log(0.01))
names = ["means3D", "scales", "rotations", "opacities", "shs"]
host = {k: cl[k].pin_memory() for k in names}
cams = [S.nerf_synthetic_camera(v, W, H).to(dev) for v in range(8)]
bg = torch.zeros(3, device=dev)
sets = [GaussianRasterizationSettings(H, W, math.tan(c.FoVx / 2), math.tan(c.FoVy / 2), bg, 1.0, c.world_view_transform,
                                      c.full_proj_transform, 3, c.camera_center, False, False) for c in cams]
dLs = [torch.randn(3, H, W, generator=torch.Generator().manual_seed(7 + i)).to(dev) for i in range(8)]
staged = MV.StagedInputs(host, dev)

def h2d_only(n=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(n):
        staged.prefetch(i & 1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / n
    print(f"H2D only: {dt*1e3:.2f} ms/step = {staged.bytes_per_step/dt/1e9:.1f} GB/s")

def views(bufs):
    t = {k: bufs[k].detach().requires_grad_() for k in names}
    loss = torch.zeros((), device=dev)
    for j in range(8):
        m2d = torch.zeros_like(t["means3D"], requires_grad=True)
        img, _ = GaussianRasterizer(sets[j])(means3D=t["means3D"], means2D=m2d, shs=t["shs"], opacities=t["opacities"],
                                             scales=t["scales"], rotations=t["rotations"])
        l = (img * dLs[j]).sum(); l.backward(); loss = loss + l.detach()
    return float(loss.cpu())

def resident(n=5):
    bufs = staged.acquire(0)
    for _ in range(2): views(bufs)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): views(bufs)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / n
    print(f"autograd module path, resident inputs: {dt*1e3:.2f} ms/step = {8/dt:.0f} frames/s")
    t = time.perf_counter()
    for _ in range(n):
        views(bufs)
    host_dt = (time.perf_counter() - t) / n
    print(f"  (host-side issue time incl. final loss read: {host_dt*1e3:.2f} ms/step)")

h2d_only(); resident(); h2d_only()
dmgs_b200.check_async()

# ---- host-side enqueue time vs device time of the autograd path, and a cProfile of the host side
import cProfile, pstats, io
bufs = staged.acquire(0)
def enqueue_only():
    t = {k: bufs[k].detach().requires_grad_() for k in names}
    for j in range(8):
        m2d = torch.zeros_like(t["means3D"], requires_grad=True)
        img, _ = GaussianRasterizer(sets[j])(means3D=t["means3D"], means2D=m2d, shs=t["shs"], opacities=t["opacities"],
                                             scales=t["scales"], rotations=t["rotations"])
        (img * dLs[j]).sum().backward()
for _ in range(2): enqueue_only()
torch.cuda.synchronize()
t0 = time.perf_counter(); enqueue_only(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"autograd path, 8 views: host enqueue {1e3*(t1-t0):.2f} ms, until device idle {1e3*(t2-t0):.2f} ms")
pr = cProfile.Profile(); pr.enable(); enqueue_only(); pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
dmgs_b200.check_async()
