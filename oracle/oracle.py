"""ctypes front end of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this
module.  Inputs and outputs are numpy arrays (float32 / int32 / uint32 / uint64 / int64).
See oracle/splat_oracle.c for the reference file:line each function follows and for the parity
status (binding + SH pinned by tests/golden; rasteriser "parity unpinned").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcParams(C.Structure):
    _fields_ = [("P", C.c_int32), ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32), ("W", C.c_int32),
                ("H", C.c_int32), ("sh_layout", C.c_int32), ("sh_act", C.c_int32), ("pad_", C.c_int32),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
                ("bg", C.c_float * 3), ("view", C.c_float * 16), ("proj", C.c_float * 16),
                ("campos", C.c_float * 3)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "splat_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _load(path):
    l = C.CDLL(path)
    l.orc_count_instances.restype = C.c_int64
    l.orc_num_threads.restype = C.c_int
    return l


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _load(build())
    return _LIB


def use_native() -> str:
    """CPU-baseline timing only (bench.py): rebuild the same source with `-O3 -march=native` ON THE
    MACHINE THAT RUNS IT (the portable liboracle.so travels to the GPU box prebuilt for x86-64-v3) and
    switch to it.  Falls back to the portable build if gcc is unavailable.  Returns the flags in use."""
    global _LIB
    out_dir = os.path.join(_HERE, "_native")
    so = os.path.join(out_dir, "liboracle_native.so")
    src = os.path.join(_HERE, "splat_oracle.c")
    flags = ["-O3", "-march=native", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math"]
    try:
        os.makedirs(out_dir, exist_ok=True)
        tmp = so + f".{os.getpid()}.tmp"
        subprocess.check_call(["gcc", *flags, "-o", tmp, src, "-lm"], stderr=subprocess.DEVNULL)
        os.replace(tmp, so)  # always rebuilt: the file may come from another machine's snapshot
        _LIB = _load(so)
        return " ".join(flags)
    except Exception:
        lib()
        return "portable build (oracle/Makefile flags); native rebuild failed"


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def make_params(P, W, H, tanfovx, tanfovy, bg, view, proj, campos, sh_degree=3, sh_coeffs=16,
                scale_modifier=1.0, sh_layout=0, sh_act=0) -> OrcParams:
    pr = OrcParams()
    pr.P, pr.W, pr.H = int(P), int(W), int(H)
    pr.sh_degree, pr.sh_coeffs, pr.sh_layout, pr.sh_act = int(sh_degree), int(sh_coeffs), int(sh_layout), int(sh_act)
    pr.tanfovx, pr.tanfovy, pr.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
    pr.bg[:] = [float(v) for v in np.asarray(bg).reshape(-1)]
    pr.view[:] = [float(v) for v in np.asarray(view, dtype=np.float32).reshape(-1)]
    pr.proj[:] = [float(v) for v in np.asarray(proj, dtype=np.float32).reshape(-1)]
    pr.campos[:] = [float(v) for v in np.asarray(campos, dtype=np.float32).reshape(-1)]
    return pr


def exp_array(x):
    x = _f32(x)
    y = np.empty_like(x)
    lib().orc_exp_array(_p(x), _p(y), C.c_int64(x.size))
    return y


def preprocess(pr: OrcParams, means3D, opacities, scales=None, rotations=None, cov3D_precomp=None,
               shs=None, colors_precomp=None) -> dict:
    P = pr.P
    means3D, opacities = _f32(means3D), _f32(opacities)
    scales, rotations, cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
    shs, colors_precomp = _f32(shs), _f32(colors_precomp)
    out = {"depths": np.empty(P, np.float32), "radii": np.empty(P, np.int32), "xy": np.empty((P, 2), np.float32),
           "conic_opacity": np.empty((P, 4), np.float32), "rgb": np.empty((P, 3), np.float32),
           "clamped": np.empty((P, 3), np.uint8), "cov3D": np.empty((P, 6), np.float32),
           "tiles_touched": np.empty(P, np.uint32)}
    lib().orc_preprocess(C.byref(pr), _p(means3D), _p(scales), _p(rotations), _p(cov3D_precomp), _p(opacities),
                         _p(shs), _p(colors_precomp), _p(out["depths"]), _p(out["radii"]), _p(out["xy"]),
                         _p(out["conic_opacity"]), _p(out["rgb"]), _p(out["clamped"]), _p(out["cov3D"]),
                         _p(out["tiles_touched"]))
    return out


def binning(pr: OrcParams, geom: dict) -> dict:
    R = int(lib().orc_count_instances(C.c_int32(pr.P), _p(geom["tiles_touched"])))
    gx, gy = (pr.W + 15) // 16, (pr.H + 15) // 16
    keys = np.empty(max(R, 1), np.uint64)
    vals = np.empty(max(R, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    lib().orc_binning(C.c_int32(pr.P), C.c_int32(pr.W), C.c_int32(pr.H), _p(geom["depths"]), _p(geom["xy"]),
                      _p(geom["radii"]), _p(geom["tiles_touched"]), C.c_int64(R), _p(keys), _p(vals), _p(ranges))
    return {"R": R, "keys": keys[:R], "vals": vals[:R], "ranges": ranges}


def blend_forward(pr: OrcParams, geom: dict, bins: dict) -> dict:
    H, W = pr.H, pr.W
    out = {"color": np.empty((3, H, W), np.float32), "final_T": np.empty((H, W), np.float32),
           "n_contrib": np.empty((H, W), np.uint32)}
    vals = bins["vals"] if bins["R"] > 0 else np.zeros(1, np.uint32)
    lib().orc_blend_forward(C.byref(pr), _p(bins["ranges"]), _p(vals), _p(geom["xy"]), _p(geom["conic_opacity"]),
                            _p(geom["rgb"]), _p(out["color"]), _p(out["final_T"]), _p(out["n_contrib"]))
    return out


def blend_backward(pr: OrcParams, geom: dict, bins: dict, img: dict, dL_dpix, abs_sums: bool = False) -> dict:
    """abs_sums: also return "abs9" [P,9] (order mean2D.xy, conic.abc, opacity, colour.rgb): the sum of the magnitudes
    of the terms of every gradient sum (its conditioning; see orc_blend_backward)."""
    P = pr.P
    dL_dpix = _f32(dL_dpix)
    out = {"dL_dmean2D": np.empty((P, 2), np.float32), "dL_dconic": np.empty((P, 3), np.float32),
           "dL_dopacity": np.empty(P, np.float32), "dL_dcolor": np.empty((P, 3), np.float32)}
    if abs_sums:
        out["abs9"] = np.empty((P, 9), np.float32)
    vals = bins["vals"] if bins["R"] > 0 else np.zeros(1, np.uint32)
    lib().orc_blend_backward(C.byref(pr), _p(bins["ranges"]), _p(vals), _p(geom["xy"]), _p(geom["conic_opacity"]),
                             _p(geom["rgb"]), _p(img["final_T"]), _p(img["n_contrib"]), _p(dL_dpix),
                             _p(out["dL_dmean2D"]), _p(out["dL_dconic"]), _p(out["dL_dopacity"]), _p(out["dL_dcolor"]),
                             _p(out.get("abs9")))
    return out


def preprocess_backward(pr: OrcParams, geom: dict, bgrad: dict, means3D, scales=None, rotations=None,
                        shs=None, precomp_color=False) -> dict:
    P, M = pr.P, pr.sh_coeffs
    means3D, scales, rotations, shs = _f32(means3D), _f32(scales), _f32(rotations), _f32(shs)
    out = {"dL_dmeans3D": np.empty((P, 3), np.float32), "dL_dcov3D": np.empty((P, 6), np.float32)}
    if scales is not None:
        out["dL_dscales"] = np.empty((P, 3), np.float32)
        out["dL_drotations"] = np.empty((P, 4), np.float32)
    if shs is not None:
        out["dL_dshs"] = np.empty(shs.shape, np.float32)
    if precomp_color:
        out["dL_dcolors_precomp"] = np.empty((P, 3), np.float32)
    lib().orc_preprocess_backward(C.byref(pr), _p(means3D), _p(scales), _p(rotations), _p(shs), _p(geom["radii"]),
                                  _p(geom["cov3D"]), _p(geom["clamped"]), _p(geom["rgb"]), _p(bgrad["dL_dmean2D"]),
                                  _p(bgrad["dL_dconic"]), _p(bgrad["dL_dcolor"]), _p(out["dL_dmeans3D"]),
                                  _p(out["dL_dcov3D"]), _p(out.get("dL_dscales")), _p(out.get("dL_drotations")),
                                  _p(out.get("dL_dshs")), _p(out.get("dL_dcolors_precomp")))
    return out


def render_forward(pr: OrcParams, means3D, opacities, **kw) -> dict:
    geom = preprocess(pr, means3D, opacities, **kw)
    bins = binning(pr, geom)
    img = blend_forward(pr, geom, bins)
    return {"geom": geom, "bins": bins, "img": img}


def render_backward(pr: OrcParams, fwd: dict, dL_dpix, means3D, scales=None, rotations=None, shs=None,
                    precomp_color=False, abs_sums: bool = False) -> dict:
    bg = blend_backward(pr, fwd["geom"], fwd["bins"], fwd["img"], dL_dpix, abs_sums)
    pg = preprocess_backward(pr, fwd["geom"], bg, means3D, scales, rotations, shs, precomp_color)
    pg.update(bg)
    return pg


def bind_forward(verts, faces, bc, rad_base, thin_z, g, adaptive=True) -> dict:
    verts, bc = _f32(verts), _f32(bc)
    faces = np.ascontiguousarray(np.asarray(faces, dtype=np.int64))
    F, k = faces.shape[0], bc.shape[0]
    out = {"xyz": np.empty((F * k, 3), np.float32), "cov6": np.empty((F * k, 6), np.float32),
           "rot_t2w": np.empty((F, 3, 3), np.float32), "cov3D_L": np.empty((F, 3, 3), np.float32)}
    lib().orc_bind_forward(C.c_int64(F), C.c_int32(k), _p(verts), _p(faces), _p(bc), C.c_float(rad_base),
                           C.c_float(thin_z), C.c_float(g), C.c_int32(int(adaptive)), _p(out["xyz"]), _p(out["cov6"]),
                           _p(out["rot_t2w"]), _p(out["cov3D_L"]))
    return out


def bind_backward(verts, faces, bc, rad_base, thin_z, g, dL_dxyz, dL_dcov6, adaptive=True) -> dict:
    verts, bc, dL_dxyz, dL_dcov6 = _f32(verts), _f32(bc), _f32(dL_dxyz), _f32(dL_dcov6)
    faces = np.ascontiguousarray(np.asarray(faces, dtype=np.int64))
    F, k, V = faces.shape[0], bc.shape[0], verts.shape[0]
    dverts = np.empty((V, 3), np.float64)
    dg = C.c_double(0.0)
    lib().orc_bind_backward(C.c_int64(F), C.c_int32(k), C.c_int64(V), _p(verts), _p(faces), _p(bc),
                            C.c_float(rad_base), C.c_float(thin_z), C.c_float(g), C.c_int32(int(adaptive)),
                            _p(dL_dxyz), _p(dL_dcov6), _p(dverts), C.byref(dg))
    return {"dverts": dverts, "dg": dg.value}


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_threads(n: int) -> None:
    """OpenMP threads of the oracle (torchrun exports OMP_NUM_THREADS=1; the CPU baseline uses every core)."""
    lib().orc_set_threads(C.c_int(int(n)))
