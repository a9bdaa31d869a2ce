#!/usr/bin/env python
"""How far is the product's arithmetic contract from upstream's default nvcc arithmetic?

The product (and its CPU oracle) compute the forward in ONE private contract (-fmad=false, explicit
fused multiply-adds, a fixed-sequence exp) -- that makes kernel == oracle bit
for bit, but the upstream rasteriser compiled by nvcc with default flags (contraction on, libm expf,
GLM expression grouping, `ndc2Pix` in double) rounds differently, so integer decisions may flip at
1-ulp boundaries.  This script counts how often, on the BASELINE.json shapes:

  * Gaussians whose radius / tile rectangle / culling decision differs,
  * the difference in the instance count R and in the sorted (tile | depth) key multiset,
  * pixels whose contributor count n_contrib differs, and the image / final-T difference.

"Upstream arithmetic" = oracle/libupstream_arith.so (test infrastructure; the algorithm of SURVEY.md
Appendix A typed in upstream's grouping and compiled with nvcc defaults).  It is NOT the reference.

    python scripts/arith_divergence.py [--out profiles/r2_arith_divergence.json] [c1 h0 c3 ...]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dmgs_b200 import GaussianRasterizationSettings, synthetic as S  # noqa: E402
from dmgs_b200.rasterizer import rasterize_forward  # noqa: E402

SHAPES = {
    "c1": (100_000, 800, 800, "nerf", 1.3, math.log(0.01)),
    "h0": (1_000_000, 800, 800, "nerf", 1.3, math.log(0.01)),
    "c3": (1_000_000, 1245, 825, "bicycle", 3.0, math.log(0.008)),
    "c5_1080p": (1_000_000, 1920, 1080, "nerf", 1.3, math.log(0.01)),
    "small": (3000, 200, 136, "small", 1.0, math.log(0.05)),
}


def ua_lib():
    l = C.CDLL(os.path.join(ROOT, "oracle", "libupstream_arith.so"))
    return l


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def upstream_forward(lib, cam, W, H, bg, d, rgb4):
    """The forward in upstream arithmetic; binning with torch (stable sort of the 64-bit keys)."""
    dev = d["means3D"].device
    P = d["means3D"].shape[0]
    z = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)
    depths, radii, xy, co = z(P), z(P, dt=torch.int32), z(P, 2), z(P, 4)
    rect, tiles = z(P, 4, dt=torch.int32), z(P, dt=torch.int32)
    view = (C.c_float * 16)(*cam.world_view_transform.reshape(-1).tolist())
    proj = (C.c_float * 16)(*cam.full_proj_transform.reshape(-1).tolist())
    rc = lib.ua_preprocess(P, W, H, C.c_float(math.tan(cam.FoVx / 2)), C.c_float(math.tan(cam.FoVy / 2)), C.c_float(1.0),
                           view, proj, ptr(d["means3D"]), ptr(d["scales"]), ptr(d["rotations"]), None, ptr(d["opacities"]),
                           ptr(depths), ptr(radii), ptr(xy), ptr(co), ptr(rect), ptr(tiles), None)
    assert rc == 0, rc
    gx, gy = (W + 15) // 16, (H + 15) // 16
    n = tiles.long()
    idx = torch.nonzero(n > 0).squeeze(1)
    nn = n[idx]
    rep = torch.repeat_interleave(idx, nn)
    off = torch.arange(int(nn.sum()), device=dev) - torch.repeat_interleave(torch.cumsum(nn, 0) - nn, nn)
    w = (rect[:, 1] - rect[:, 0]).long()[rep]
    ty = rect[:, 2].long()[rep] + off // w
    tx = rect[:, 0].long()[rep] + off % w
    keys = ((ty * gx + tx) << 32) | (depths.view(torch.int32)[rep].long() & 0xFFFFFFFF)
    keys_sorted, order = torch.sort(keys, stable=True)
    gidx = rep[order].int().contiguous()
    tile_of = (keys_sorted >> 32)
    counts = torch.bincount(tile_of, minlength=gx * gy)
    ends = torch.cumsum(counts, 0)
    ranges = torch.stack([ends - counts, ends], 1).int().contiguous()
    color, fT, nc = z(3, H, W), z(H, W), z(H, W, dt=torch.int32)
    bgc = (C.c_float * 3)(*bg)
    rc = lib.ua_blend(W, H, bgc, ptr(ranges), ptr(gidx), ptr(xy), ptr(co), ptr(rgb4), 4, ptr(color), ptr(fT), ptr(nc), None)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return dict(depths=depths, radii=radii, xy=xy, co=co, rect=rect, tiles=tiles, keys=keys_sorted, gidx=gidx,
                color=color, final_T=fT, n_contrib=nc)


def compare(name, view=0):
    P, W, H, kind, extent, lsm = SHAPES[name]
    if kind == "small":
        cam = S.look_at_camera([2.5, 1.0, 1.2], W, H, fovx=0.9)
    else:
        cam = S.nerf_synthetic_camera(view, W, H) if kind == "nerf" else S.bicycle_camera(view, W, H)
    cl = S.random_cloud(P, seed=0 if kind != "small" else 3, extent=extent, log_scale_mean=lsm)
    d = {k: v.cuda() for k, v in cl.items()}
    bg = (0.0, 0.0, 0.0)
    rs = GaussianRasterizationSettings(H, W, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2), torch.tensor(bg, device="cuda"),
                                       1.0, cam.world_view_transform.cuda(), cam.full_proj_transform.cuda(), 3,
                                       cam.camera_center.cuda(), False, False)
    color, radii, st = rasterize_forward(rs, d["means3D"], d["opacities"], d["shs"], None, d["scales"], d["rotations"], None)
    torch.cuda.synchronize()
    g, b, im = st.geom_arrays(), st.binning_arrays(), st.image_arrays()
    ours_keys = st.sorted_keys()
    ua = upstream_forward(ua_lib(), cam, W, H, bg, d, g["rgb"].contiguous())
    vis_o, vis_u = radii > 0, ua["radii"] > 0
    both = vis_o & vis_u
    rect_o = g["rect"].int()
    bits = lambda t: t.contiguous().view(torch.int32)
    ulp = lambda a, b_: (bits(a).long() - bits(b_).long()).abs()
    xy_ulp = ulp(g["rec"][:, 0:2][both], ua["xy"][both])
    con_ulp = ulp(g["rec"][:, 2:5][both], ua["co"][:, 0:3][both])
    # key multisets: size of the symmetric difference
    ku, ko = ua["keys"], ours_keys
    cat = torch.cat([ko, ku])
    uniq, cnt = torch.unique(cat, return_counts=True)
    in_o = torch.isin(uniq, ko)
    # multiset counts per side
    co_ = torch.bincount(torch.searchsorted(uniq, ko), minlength=uniq.numel())
    cu_ = torch.bincount(torch.searchsorted(uniq, ku), minlength=uniq.numel())
    key_symdiff = int((co_ - cu_).abs().sum())
    nc_o, nc_u = im["n_contrib"], ua["n_contrib"]
    out = {
        "shape": name, "P": P, "W": W, "H": H, "view": view,
        "visible_ours": int(vis_o.sum()), "visible_upstream_arith": int(vis_u.sum()),
        "culling_decision_differs": int((vis_o != vis_u).sum()),
        "radius_differs": int((radii[both] != ua["radii"][both]).sum()),
        "tile_rect_differs": int((rect_o[both] != ua["rect"][both]).any(1).sum()),
        "tiles_touched_differs": int((g["tiles_touched"][both] != ua["tiles"][both]).sum()),
        "depth_bits_differ": int((bits(g["depths"])[both] != bits(ua["depths"])[both]).sum()),
        "xy_bits_differ": int((xy_ulp > 0).any(1).sum()), "xy_max_ulp": int(xy_ulp.max()) if xy_ulp.numel() else 0,
        "conic_bits_differ": int((con_ulp > 0).any(1).sum()), "conic_max_ulp": int(con_ulp.max()) if con_ulp.numel() else 0,
        "R_ours": int(st.num_rendered), "R_upstream_arith": int(ua["keys"].numel()),
        "sorted_key_multiset_symmetric_difference": key_symdiff,
        "sorted_list_identical": bool(ko.numel() == ku.numel() and torch.equal(ko, ku) and
                                      torch.equal(b["gidx"], ua["gidx"])),
        "pixels": W * H,
        "n_contrib_differs_pixels": int((nc_o != nc_u).sum()),
        "n_contrib_max_abs_diff": int((nc_o.long() - nc_u.long()).abs().max()),
        "final_T_max_abs_diff": float((im["final_T"] - ua["final_T"]).abs().max()),
        "image_max_abs_diff": float((color - ua["color"]).abs().max()),
        "image_pixels_over_1e-5": int(((color - ua["color"]).abs().amax(0) > 1e-5).sum()),
    }
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shapes", nargs="*", default=["c1", "h0", "c3"])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    res = {"what": "product arithmetic contract (-fmad=false, explicit fma, dmgs_exp) vs the same algorithm in upstream's "
                   "expression grouping under nvcc default arithmetic (-fmad=true, libm expf, ndc2Pix in double); "
                   "oracle/upstream_arith.cu; counts per frame",
           "frames": [compare(s) for s in args.shapes]}
    txt = json.dumps(res, indent=1)
    print(txt)
    if args.out:
        os.makedirs(os.path.dirname(os.path.join(ROOT, args.out)), exist_ok=True)
        with open(os.path.join(ROOT, args.out), "w") as fh:
            fh.write(txt + "\n")


if __name__ == "__main__":
    main()
