"""The C-ABI boundary on a machine without a GPU: libdmgs_raster.so loads, exports every function
include/dmgs_raster.h declares, the ctypes binding covers all of them, the host-only size functions
answer, and every product entry point refuses CPU tensors (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dmgs_raster.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)  # prototypes only, not names mentioned in comments
    return sorted(set(re.findall(r"\b(dmgs_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def so_path():
    from dmgs_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "dmgs_b200", "csrc"), "-j8", "-s"])
    return _lib.LIB_PATH


def test_library_exports_every_declared_symbol(so_path):
    names = declared_functions()
    assert len(names) >= 28
    lib = C.CDLL(so_path)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in dmgs_raster.h but not exported: {missing}"


def test_ctypes_binding_covers_the_header(so_path):
    from dmgs_b200 import _lib
    declared = set(declared_functions())
    bound = set(_lib.EXPORTS)
    assert declared - bound == set(), f"no ctypes signature for {sorted(declared - bound)}"
    assert bound - declared <= {"dmgs_launch_count"}, f"bound but undeclared: {sorted(bound - declared)}"
    l = _lib.lib()
    assert l.dmgs_abi_version() == 1
    # host-only functions (no CUDA calls): sizes are positive, 256-byte aligned and monotone
    assert l.dmgs_geom_bytes(1000) % 256 == 0 and l.dmgs_geom_bytes(2000) > l.dmgs_geom_bytes(1000) > 0
    assert l.dmgs_binning_bytes(1000, 50000, 800, 800) > 0 and l.dmgs_image_bytes(800, 800) >= 800 * 800 * 8
    assert l.dmgs_backward_scratch_bytes(1000) >= 1000 * 48
    assert l.dmgs_l1_ssim_scratch_bytes(3, 800, 800) >= 3 * 3 * 800 * 800 * 4
    assert l.dmgs_frustum_scratch_bytes(10 ** 6) >= (10 ** 6 // 256) * 4


def test_no_cpu_fallback():
    from dmgs_b200 import GaussianRasterizationSettings, GaussianRasterizer, frustum, loss_utils
    from dmgs_b200.binding import bind_faces
    from dmgs_b200.optim import FusedAdam
    eye = torch.eye(4)
    rs = GaussianRasterizationSettings(32, 32, 0.5, 0.5, torch.zeros(3), 1.0, eye, eye, 0, torch.zeros(3), False, False)
    P = 4
    with pytest.raises(RuntimeError, match="no CPU path"):
        GaussianRasterizer(rs)(means3D=torch.zeros(P, 3), means2D=torch.zeros(P, 3), opacities=torch.ones(P, 1),
                               colors_precomp=torch.ones(P, 3), scales=torch.ones(P, 3), rotations=torch.ones(P, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        loss_utils.l1_ssim_loss(torch.zeros(3, 8, 8), torch.zeros(3, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        frustum.in_frustum(eye, torch.zeros(5, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        frustum.cull_faces(eye, torch.zeros(5, 3), torch.zeros(2, 3, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        FusedAdam([{"params": [torch.zeros(4)], "lr": 0.1, "name": "x"}])
    with pytest.raises(RuntimeError, match="no CPU path"):
        bind_faces(torch.zeros(3, 3), torch.zeros(1, 3, dtype=torch.int64), torch.ones(1, 3) / 3, 1.0, 1.0, None)


def test_scheduling_knobs_are_host_only_and_leave_the_layouts_alone(so_path):
    """dmgs_set_place_smem_kb / dmgs_set_blend_residency only steer how kernels are launched: argument ranges are
    checked, and the caller-owned buffer sizes (which a caller may have computed before changing a knob) do not move."""
    from dmgs_b200 import _lib
    l = _lib.lib()
    shapes = [(1_000_000, 7_400_000, 800, 800), (491_520, 900_000, 800, 800), (3_000_000, 4_000_000, 1245, 825),
              (100_000, 2_000_000, 1920, 1080), (10, 40, 64, 48)]
    try:
        ref = [(l.dmgs_binning_bytes(*s), l.dmgs_geom_bytes(s[0]), l.dmgs_image_bytes(s[2], s[3]),
                l.dmgs_backward_scratch_bytes(s[0])) for s in shapes]
        for kb in (64, 100, 128, 160, 200):
            assert l.dmgs_set_place_smem_kb(kb) == 0
            for res in ((8, 8), (6, 6), (1, 3)):
                assert l.dmgs_set_blend_residency(*res) == 0
                got = [(l.dmgs_binning_bytes(*s), l.dmgs_geom_bytes(s[0]), l.dmgs_image_bytes(s[2], s[3]),
                        l.dmgs_backward_scratch_bytes(s[0])) for s in shapes]
                assert got == ref
        for bad in (0, 32, 63, 201, 1024, -1):
            assert l.dmgs_set_place_smem_kb(bad) == -7
        for bad in ((9, 1), (1, 9), (-1, 1), (1, -1)):
            assert l.dmgs_set_blend_residency(*bad) == -7
        assert l.dmgs_set_blend_residency(0, 0) == 0  # 0 leaves a value unchanged
    finally:
        l.dmgs_set_place_smem_kb(200)
        l.dmgs_set_blend_residency(8, 8)
    # the scratch buffer carries the backward's square counter behind the P x 12 sums
    assert l.dmgs_backward_scratch_bytes(1000) >= 1000 * 48 + 4
