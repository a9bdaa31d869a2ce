"""Runs the reference's UNMODIFIED gaussian_renderer/__init__.py against this repository's drop-in
`diff_gaussian_rasterization` package on the CPU and records what reaches the rasteriser
(SURVEY.md section 4 iv).  Writes tests/golden/ref_glue.npz.

    python tests/golden/make_golden_glue.py        (build container only: needs /root/reference)

How: the reference file is exec()'d as is (`import diff_gaussian_rasterization` resolves to the shim at the repo
root, `scene.gaussian_model` -- whose own imports need simple_knn / plyfile -- is replaced by an empty stand-in
because the glue only uses it as a type annotation, `utils.sh_utils` is the reference's own module).  The glue then
constructs the settings tuple and calls OUR `GaussianRasterizer.forward` (argument validation included); only the
C call underneath (`rasterize_gaussians`) is replaced by a recorder, because this container has no GPU.  Three CPU
shims keep the glue's hard-coded `.cuda()` / `device="cuda"` from raising.  Nothing of the reference is copied
into the repository except the recorded numbers.
"""
import math
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(OUT, "..", ".."))


class Recorder:
    def __init__(self):
        self.calls = []

    def __call__(self, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                 raster_settings, sh_layout=0, sh_activation=0, holder=None):
        self.calls.append(dict(means3D=means3D, means2D=means2D, shs=sh, colors_precomp=colors_precomp,
                               opacities=opacities, scales=scales, rotations=rotations, cov3D_precomp=cov3Ds_precomp,
                               settings=raster_settings, sh_layout=sh_layout, sh_activation=sh_activation))
        P = means3D.shape[0]
        H, W = int(raster_settings.image_height), int(raster_settings.image_width)
        return torch.zeros(3, H, W), torch.zeros(P, dtype=torch.int32)


def load_reference_glue(recorder):
    """-> namespace of the exec()'d reference gaussian_renderer/__init__.py wired to the shim + recorder."""
    for p in (ROOT, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import diff_gaussian_rasterization  # noqa: F401  (the shim: re-exports dmgs_b200.rasterizer)
    import dmgs_b200.rasterizer as R
    R.rasterize_gaussians = recorder
    if "scene.gaussian_model" not in sys.modules:
        scene = types.ModuleType("scene")
        gm = types.ModuleType("scene.gaussian_model")
        gm.GaussianModel = type("GaussianModel", (), {})
        scene.gaussian_model = gm
        sys.modules.setdefault("scene", scene)
        sys.modules["scene.gaussian_model"] = gm
    torch.Tensor.cuda = lambda self, *a, **k: self
    if not getattr(torch.zeros_like, "_cpu_shim", False):
        orig = torch.zeros_like

        def zeros_like(*a, **k):
            k.pop("device", None)
            return orig(*a, **k)
        zeros_like._cpu_shim = True
        torch.zeros_like = zeros_like
    ns = {"__name__": "gaussian_renderer"}
    with open(os.path.join(REF, "gaussian_renderer", "__init__.py")) as fh:
        exec(compile(fh.read(), "reference:gaussian_renderer/__init__.py", "exec"), ns)
    return ns


def scene(P=1500, seed=11):
    sys.path.insert(0, ROOT)
    from dmgs_b200 import synthetic as S
    cl = S.random_cloud(P, seed=seed, extent=1.0, log_scale_mean=math.log(0.04))
    g = torch.Generator().manual_seed(seed)
    cov = torch.rand(P, 6, generator=g) * 1e-3
    cov[:, [0, 3, 5]] += 2e-3  # diagonally dominant: positive definite
    feats_p3m = (torch.randn(P, 3, 16, generator=g) * 0.3).contiguous()
    cam = S.nerf_synthetic_camera(2, 200, 136)
    return cl, cov, feats_p3m, cam


class DuckModel:
    """What gaussian_renderer.render reads from a GaussianModel (scene/gaussian_model.py properties)."""

    def __init__(self, cl, cov):
        self._cl, self._cov = cl, cov
        self.active_sh_degree, self.max_sh_degree = 3, 3

    get_xyz = property(lambda s: s._cl["means3D"])
    get_opacity = property(lambda s: s._cl["opacities"])
    get_scaling = property(lambda s: s._cl["scales"])
    get_rotation = property(lambda s: s._cl["rotations"])
    get_features = property(lambda s: s._cl["shs"])

    def get_covariance(self, scaling_modifier=1.0):
        return self._cov * scaling_modifier


CASES = [("render", dict(compute_cov3D_python=False, convert_SHs_python=False)),
         ("render", dict(compute_cov3D_python=True, convert_SHs_python=False)),
         ("render", dict(compute_cov3D_python=False, convert_SHs_python=True)),
         ("render", dict(compute_cov3D_python=False, convert_SHs_python=False, override=True)),
         ("render_dyn", dict(compute_cov3D_python=True, convert_SHs_python=True)),
         ("render_dyn", dict(compute_cov3D_python=True, convert_SHs_python=True, override=True))]


def run_cases(ns, rec):
    """Calls the glue functions in `ns` (reference or mirror) for every case; -> list of recorded calls."""
    cl, cov, feats, cam = scene()
    bg = torch.tensor([0.1, 0.2, 0.3])
    override = torch.rand(cl["means3D"].shape[0], 3, generator=torch.Generator().manual_seed(3))
    out = []
    for fn, flags in CASES:
        flags = dict(flags)
        ov = override if flags.pop("override", False) else None
        pipe = SimpleNamespace(debug=False, **flags)
        n0 = len(rec.calls)
        if fn == "render":
            res = ns["render"](cam, DuckModel(cl, cov), pipe, bg, 1.0, ov)
        else:
            gs = dict(xyz=cl["means3D"], opacity=cl["opacities"], covariance=cov, features=feats, active_sh_degree=3,
                      max_sh_degree=3)
            res = ns["render_dyn"](cam, gs, pipe, bg, 1.0, ov)
        assert len(rec.calls) == n0 + 1, "the glue must reach the rasteriser exactly once"
        assert set(res) == {"render", "viewspace_points", "visibility_filter", "radii"}
        assert res["viewspace_points"] is rec.calls[-1]["means2D"]
        out.append(rec.calls[-1])
    return out


def main():
    rec = Recorder()
    ns = load_reference_glue(rec)
    calls = run_cases(ns, rec)
    save = {}
    for i, c in enumerate(calls):
        for k in ("means3D", "shs", "colors_precomp", "opacities", "scales", "rotations", "cov3D_precomp"):
            if c[k] is not None:
                save[f"c{i}_{k}"] = c[k].detach().numpy()
        s = c["settings"]
        save[f"c{i}_settings"] = np.array([s.image_height, s.image_width, s.tanfovx, s.tanfovy, s.scale_modifier,
                                           s.sh_degree, float(s.prefiltered), float(s.debug)], np.float64)
        save[f"c{i}_bg"], save[f"c{i}_view"] = s.bg.numpy(), s.viewmatrix.numpy()
        save[f"c{i}_proj"], save[f"c{i}_campos"] = s.projmatrix.numpy(), s.campos.numpy()
    np.savez_compressed(os.path.join(OUT, "ref_glue.npz"), **save)
    print("wrote ref_glue.npz:", len(calls), "calls,", sum(v.nbytes for v in save.values()) // 1024, "KiB")


def load_mirror(recorder):
    """-> {"render", "render_dyn"} of dmgs_b200.renderer wired to the same recorder."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import dmgs_b200.rasterizer as R
    import dmgs_b200.renderer as M
    R.rasterize_gaussians = recorder
    return {"render": M.render, "render_dyn": M.render_dyn}


def compare_calls(ref_calls, mir_calls):
    """The mirror must hand the rasteriser what the reference glue hands it; where the mirror folds the
    reference's python SH evaluation into the kernel (shs + sigmoid flag instead of colors_precomp) the
    reference's colours must be what that SH evaluation gives.  -> list of human-readable findings."""
    from oracle import torch_oracle as TO
    notes = []
    for i, (r, m) in enumerate(zip(ref_calls, mir_calls)):
        for k in ("means3D", "opacities", "scales", "rotations", "cov3D_precomp"):
            assert (r[k] is None) == (m[k] is None), f"case {i}: {k} None-ness differs"
            if r[k] is not None:
                assert torch.equal(torch.as_tensor(r[k]), torch.as_tensor(m[k])), f"case {i}: {k} differs"
        rs, ms = r["settings"], m["settings"]
        for f in ("image_height", "image_width", "tanfovx", "tanfovy", "scale_modifier", "sh_degree", "prefiltered", "debug"):
            assert getattr(rs, f) == getattr(ms, f), f"case {i}: settings.{f} {getattr(rs, f)} vs {getattr(ms, f)}"
        for f in ("bg", "viewmatrix", "projmatrix", "campos"):
            assert torch.equal(torch.as_tensor(getattr(rs, f)), torch.as_tensor(getattr(ms, f))), f"case {i}: settings.{f}"
        if r["colors_precomp"] is not None and m["colors_precomp"] is None:
            assert m["sh_activation"] == 1, f"case {i}: the mirror must select the sigmoid activation"
            feats = torch.as_tensor(m["shs"])
            pm3 = feats if m["sh_layout"] == 0 else feats.transpose(1, 2)
            col = TO.eval_sh_colors(int(rs.sh_degree), pm3.double(), torch.as_tensor(m["means3D"]).double(),
                                    torch.as_tensor(ms.campos).double(), 1)
            err = (col - torch.as_tensor(r["colors_precomp"]).double()).abs().max().item()
            assert err < 2e-6, f"case {i}: sigmoid(eval_sh) of the mirror's shs differs from the reference's colours by {err}"
            notes.append(f"case {i}: python SH folded into the kernel (max colour difference {err:.1e})")
        else:
            for k in ("shs", "colors_precomp"):
                assert (r[k] is None) == (m[k] is None), f"case {i}: {k} None-ness differs"
                if r[k] is not None:
                    assert torch.equal(torch.as_tensor(r[k]), torch.as_tensor(m[k])), f"case {i}: {k} differs"
            assert m["sh_activation"] == 0, f"case {i}: clamp activation expected"
            notes.append(f"case {i}: identical rasteriser arguments")
    return notes


def check():
    """Live: reference glue vs mirror vs the committed fixture.  Exit status 0 = all equal."""
    rec = Recorder()
    ref_calls = run_cases(load_reference_glue(rec), rec)
    mir_calls = run_cases(load_mirror(rec), rec)
    for n in compare_calls(ref_calls, mir_calls):
        print(n)
    gold = np.load(os.path.join(OUT, "ref_glue.npz"))
    for i, c in enumerate(ref_calls):
        for k in ("means3D", "shs", "colors_precomp", "opacities", "scales", "rotations", "cov3D_precomp"):
            key = f"c{i}_{k}"
            assert (c[k] is not None) == (key in gold.files), f"fixture out of date: {key}"
            if c[k] is not None:
                assert np.array_equal(gold[key], c[k].detach().numpy()), f"fixture out of date: {key}"
    print("fixture up to date")


if __name__ == "__main__":
    check() if "--check" in sys.argv else main()
