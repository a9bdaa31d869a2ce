"""Runs every kernel of the library a few times at full size (for one ncu --set full capture of the rows that the
H0 bench does not launch): loss, frustum cull, Adam, binding, stage-3, deferred-SH expand, peer kernels excepted."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dmgs_b200
from dmgs_b200 import frustum as FR, loss_utils as LU, multiview as MV, synthetic as S
from dmgs_b200.binding import bind_faces, bind_frame, stage3_scales_rotations
from dmgs_b200.optim import FusedAdam

dev = torch.device("cuda", 0)
dmgs_b200.configure(async_binning=True)
REP = int(os.environ.get("REP", "3"))
# loss, 800x800
gt = torch.rand(3, 800, 800, device=dev); img = (gt + 0.2 * torch.randn_like(gt)).clamp(0, 1)
for _ in range(REP):
    LU.l1_ssim_loss_and_grad(img, gt, 0.2, need_loss=False)
# frustum, 1.3 M faces
verts, faces = S.jittered_sphere_mesh(1_000_000, seed=1, jitter=0.05)
verts, faces = verts.to(dev), faces.to(dev)
cam = S.look_at_camera([0.8, 0.3, 1.1], 1245, 825, fovx=0.9)
for _ in range(REP):
    FR.cull_faces(cam.full_proj_transform.to(dev), verts, faces, 1)
# binding stage 2 + stage 3, k = 6
bc, rad = S.barycentric_layout(6)
v = verts.clone().requires_grad_(); sf = torch.tensor([math.atanh(0.5)], device=dev, requires_grad=True)
for _ in range(REP):
    xyz, cov = bind_faces(v, faces, bc.to(dev), rad, 4.43e-6, sf)
    (xyz.sum() + cov.sum() * 1e3).backward()
P3 = faces.shape[0] * 6
r2 = torch.randn(P3, 2, device=dev, requires_grad=True); s2 = (torch.randn(P3, 2, device=dev) * 0.3 - 3).requires_grad_()
for _ in range(REP):
    xyz, rot = bind_frame(v, faces, bc.to(dev))
    sc, q = stage3_scales_rotations(rot, s2, r2, 4.43e-6)
    (xyz.sum() + sc.sum() + q.sum()).backward()
# Adam on the flat buffer, 1 M Gaussians
P = 1_000_000
buf = MV.FlatGradBuffer(P, MV.RASTER_WIDTHS_SH, dev)
names = [k for k in MV.RASTER_WIDTHS_SH if k != "means2D"]
params = {k: torch.randn(P, *MV.RASTER_WIDTHS_SH[k], device=dev) for k in names}
opt = FusedAdam([{"params": [params[k]], "lr": 1e-3, "name": k} for k in names], lr=0.0, eps=1e-15)
buf.flat.normal_()
for _ in range(REP):
    opt.step(grads=buf.views, grad_scale=0.125, zero_grad=False)
torch.cuda.synchronize()
print("done")
