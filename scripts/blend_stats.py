"""DIAGNOSTIC: what the blend kernels' warp-rectangle culling does on the bench workloads (scripts/blend_stats.cu).
For each workload renders view 0 and counts, for warp rectangles 8x4 / 8x8 / 16x16: (rectangle, entry) pairs in the
backward's range, pairs surviving the cull, surviving pairs with a contributing pixel, contributing (pixel, entry)
pairs.  python scripts/blend_stats.py [h0 c1 c3] [--out profiles/r2_blend_stats.json]"""
import ctypes as C
import json, math, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import WORKLOADS, make_camera
from dmgs_b200 import GaussianRasterizationSettings, synthetic as S
from dmgs_b200.rasterizer import rasterize_forward

so = os.path.join(ROOT, "scripts", "_blend_stats.so")
src = os.path.join(ROOT, "scripts", "blend_stats.cu")
if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false",
                           "-shared", "-Xcompiler", "-fPIC", "-o", so, src])
lib = C.CDLL(so)
lib.blend_stats.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 6
args = [a for a in sys.argv[1:] if not a.startswith("--")]
out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
if out_path:
    args = [a for a in args if a != out_path]
dev = torch.device("cuda", 0)
res = {}
for name in args or ["h0"]:
    if name == "c2":  # mesh-bound stage-2 cloud (scripts/bench_configs.py c2): ~490 k small Gaussians on a sphere mesh
        from dmgs_b200.binding import bind_faces
        m = S.mesh_bound_inputs(50_000, 6, seed=1)
        W = H = 800
        c = S.nerf_synthetic_camera(0, W, H).to(dev)
        xyz, cov = bind_faces(m["verts"].to(dev), m["faces"].to(dev), m["bc"].to(dev), m["rad_base"],
                              m["spatial_lr_scale"] * 1e-6, torch.tensor([m["scale_factor"]], device=dev), 2.0)
        P = xyz.shape[0]
        st_ = GaussianRasterizationSettings(H, W, math.tan(c.FoVx / 2), math.tan(c.FoVy / 2), torch.ones(3, device=dev), 1.0,
                                            c.world_view_transform, c.full_proj_transform, 3, c.camera_center, False, False)
        color, radii, st = rasterize_forward(st_, xyz.detach(), torch.full((P, 1), 0.9999, device=dev), m["features"].to(dev),
                                             None, None, None, cov.detach(), sh_layout=1, sh_activation=1)
    else:
        P, W, H, kind, extent, lsm = WORKLOADS[name]
        cl = S.random_cloud(P, seed=0, extent=extent, log_scale_mean=lsm)
        d = {k: v.to(dev) for k, v in cl.items()}
        c = make_camera(kind, 0, W, H).to(dev)
        st_ = GaussianRasterizationSettings(H, W, math.tan(c.FoVx / 2), math.tan(c.FoVy / 2), torch.zeros(3, device=dev), 1.0,
                                            c.world_view_transform, c.full_proj_transform, 3, c.camera_center, False, False)
        color, radii, st = rasterize_forward(st_, d["means3D"], d["opacities"], d["shs"], None, d["scales"], d["rotations"], None)
    g, b, im = st.geom_arrays(), st.binning_arrays(), st.image_arrays()
    out = torch.zeros(24, dtype=torch.int64, device=dev)
    rc = lib.blend_stats(W, H, b["ranges"].data_ptr(), b["gidx"].data_ptr(), g["rec"].data_ptr(),
                         im["n_contrib"].data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0, rc
    o = out.cpu().tolist()
    R = st.num_rendered
    row = {"R": R, "consumed_entries": o[16], "consumed_frac": o[16] / max(R, 1)}
    for s, shape in enumerate(("8x4", "8x8", "16x16")):
        pairs, surv, surv_hit, hits, anyhit, rects, groups = o[8 * s:8 * s + 7]
        px = {"8x4": 32, "8x8": 64, "16x16": 256}[shape]
        row[shape] = {"pairs": pairs, "survivors": surv, "survivors_with_hit": surv_hit, "pixel_hits": hits,
                      "pairs_with_hit_ignoring_cull": anyhit, "rects_with_work": rects, "groups32": groups,
                      "survive_frac": surv / max(pairs, 1), "hit_frac_of_survivors": surv_hit / max(surv, 1),
                      "lanes_hit_per_surviving_pair": hits / max(surv, 1) / px,
                      "lanes_hit_per_hit_pair": hits / max(surv_hit, 1) / px}
        assert surv_hit == anyhit, "cull_rect dropped a contributing pair"
    rg = b["ranges"].long()
    lens = (rg[:, 1] - rg[:, 0]).float()
    row["list_len"] = {"mean": lens.mean().item(), "p50": lens.median().item(), "p99": lens.quantile(0.99).item(),
                       "max": lens.max().item()}
    nc = im["n_contrib"].float()
    row["n_contrib"] = {"mean": nc.mean().item(), "p99": nc.flatten().quantile(0.99).item(), "max": nc.max().item()}
    res[name] = row
    print(name, json.dumps(row))
if out_path:
    with open(out_path, "w") as fh:
        json.dump(res, fh, indent=1)
