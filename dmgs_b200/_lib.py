"""ctypes binding of libdmgs_raster.so (the C ABI in include/dmgs_raster.h).

There is NO CPU or PyTorch fallback: if the shared library is missing this module raises, and
every entry point needs CUDA tensors.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C dmgs_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DMGS_RASTER_LIB: another build of the same library (kernel-variant experiments, scripts/variants.sh)
LIB_PATH = os.environ.get("DMGS_RASTER_LIB") or os.path.join(_HERE, "libdmgs_raster.so")


class DmgsParams(C.Structure):
    """struct dmgs_params (include/dmgs_raster.h)."""

    _fields_ = [("P", C.c_int32), ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32), ("image_width", C.c_int32),
                ("image_height", C.c_int32), ("sh_layout", C.c_int32), ("sh_activation", C.c_int32),
                ("debug", C.c_int32), ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
                ("bg", C.c_float * 3), ("viewmatrix", C.c_float * 16), ("projmatrix", C.c_float * 16),
                ("campos", C.c_float * 3)]


class AdamSegment(C.Structure):
    """struct dmgs_adam_segment (include/dmgs_raster.h)."""

    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64), ("lr", C.c_double), ("lr_hi", C.c_double), ("period", C.c_int32), ("split", C.c_int32)]


class AdamXSegment(C.Structure):
    """struct dmgs_adam_xsegment (include/dmgs_raster.h)."""

    _fields_ = [("grad_offset", C.c_int64), ("param_offset", C.c_int64), ("n", C.c_int64), ("exp_avg_shard", C.c_void_p),
                ("exp_avg_sq_shard", C.c_void_p), ("lr", C.c_double), ("lr_hi", C.c_double), ("period", C.c_int32),
                ("split", C.c_int32)]


_lib = None

_vp, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_SIGS = {
    "dmgs_geom_bytes": (C.c_size_t, [_i32]),
    "dmgs_binning_bytes": (C.c_size_t, [_i32, _i64, _i32, _i32]),
    "dmgs_image_bytes": (C.c_size_t, [_i32, _i32]),
    "dmgs_backward_scratch_bytes": (C.c_size_t, [_i32]),
    "dmgs_preprocess_forward": (C.c_int, [C.POINTER(DmgsParams)] + [_vp] * 11),
    "dmgs_bin_forward": (C.c_int, [C.POINTER(DmgsParams), _vp, _i64, _vp, _vp]),
    "dmgs_bin_forward_async": (C.c_int, [C.POINTER(DmgsParams), _vp, _i64, _vp, _vp, _vp]),
    "dmgs_blend_forward": (C.c_int, [C.POINTER(DmgsParams), _vp, _vp, _i64, _vp, _vp, _vp]),
    "dmgs_backward": (C.c_int, [C.POINTER(DmgsParams)] + [_vp] * 9 + [_i64] + [_vp] * 11),
    "dmgs_blend_backward": (C.c_int, [C.POINTER(DmgsParams), _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "dmgs_preprocess_backward": (C.c_int, [C.POINTER(DmgsParams)] + [_vp] * 16 + [_i32, _vp]),
    "dmgs_launch_count": (C.c_uint64, []),
    "dmgs_set_blend_residency": (C.c_int, [_i32, _i32]),
    "dmgs_set_place_smem_kb": (C.c_int, [_i32]),
    "dmgs_mark_visible": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp]),
    "dmgs_bind_forward": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _f, _f, _vp, _i32, _vp, _vp, _vp, _vp]),
    "dmgs_bind_backward": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _f, _f, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dmgs_preprocess_forward_bound": (C.c_int, [C.POINTER(DmgsParams), _i64, _i32, _vp, _vp, _vp, _f, _f, _vp, _i32] + [_vp] * 8),
    "dmgs_preprocess_backward_bound": (C.c_int, [C.POINTER(DmgsParams), _i64, _i32, _vp, _vp, _vp, _f, _f, _vp, _i32]
                                       + [_vp] * 10 + [_i32, _vp]),
    "dmgs_sh_grad_expand": (C.c_int, [_i32, _i32, _i32, _i32, _i32, C.POINTER(_f), _vp, _vp, _vp, _i64, _vp, _vp, _i32, _vp]),
    "dmgs_stage3_forward": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp]),
    "dmgs_stage3_backward": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dmgs_l1_ssim_scratch_bytes": (C.c_size_t, [_i32, _i32, _i32]),
    "dmgs_l1_ssim_forward": (C.c_int, [_i32, _i32, _i32, C.POINTER(_f), _vp, _vp, _vp, _vp, _vp]),
    "dmgs_l1_ssim_backward": (C.c_int, [_i32, _i32, _i32, C.POINTER(_f), _vp, _vp, _vp, _vp, _vp, _vp]),
    "dmgs_frustum_scratch_bytes": (C.c_size_t, [_i64]),
    "dmgs_in_frustum": (C.c_int, [_i64, C.POINTER(_f), _f, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dmgs_adam_step": (C.c_int, [_i32, C.POINTER(AdamSegment), C.c_double, C.c_double, C.c_double, _i64, _f, _i32, _vp]),
    "dmgs_texture_grid_params": (_i64, []),
    "dmgs_texture_cast_params": (C.c_int, [_i64, _vp, _vp, _vp]),
    "dmgs_texture_forward": (C.c_int, [_i64, _i32, C.POINTER(_f)] + [_vp] * 8),
    "dmgs_texture_backward": (C.c_int, [_i64, _i32, C.POINTER(_f)] + [_vp] * 7 + [_f] + [_vp] * 7),
    "dmgs_texture_backward_scratch_bytes": (C.c_size_t, [_i64]),
    "dmgs_adam_exchange_shard": (C.c_int, [_i64, _i32, _i32, C.POINTER(_i64), C.POINTER(_i64)]),
    "dmgs_adam_exchange_peer": (C.c_int, [_i32, _i32, _i32, C.POINTER(AdamXSegment), C.POINTER(C.c_void_p), _vp,
                                          C.POINTER(C.c_void_p), _vp, C.c_double, C.c_double, C.c_double, _i64, _f, _vp]),
    "dmgs_allreduce_peer": (C.c_int, [_i64, _i32, _i32, C.POINTER(C.c_void_p), _vp, _f, _vp]),
    "dmgs_geom_layout": (C.c_int, [_i32, C.POINTER(_i64)]),
    "dmgs_binning_layout": (C.c_int, [_i32, _i64, _i32, _i32, C.POINTER(_i64)]),
    "dmgs_image_layout": (C.c_int, [_i32, _i32, C.POINTER(_i64)]),
    "dmgs_sorted_keys": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp]),
    "dmgs_exp_array": (C.c_int, [_vp, _vp, _i64, _vp]),
    "dmgs_last_error": (C.c_char_p, []),
    "dmgs_abi_version": (C.c_int, []),
}
EXPORTS = tuple(_SIGS)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run __graft_entry__.build() or `make -C dmgs_b200/csrc`). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().dmgs_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def ptr(t):
    """Device pointer of a (contiguous) tensor, or NULL for None."""
    return None if t is None else C.c_void_p(t.data_ptr())
