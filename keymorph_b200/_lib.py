"""ctypes binding of libkm_b200.so (C ABI declared in include/km_b200.h).

There is no CPU path behind this module: if the shared library is missing or a CUDA device is not
present the calls fail loudly.  Nothing under oracle/ is ever imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libkm_b200.so")

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_f = C.c_float
_d = C.c_double
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/km_b200.h one to one
SIGNATURES = {
    "km_version": (_i, []),
    "km_last_error": (C.c_char_p, []),
    "km_sm_count": (_i, []),
    "km_set_option": (_i, [_i, _i]),
    "km_operand_is_fp16": (_i, []),
    "km_grid_sample3d": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "km_flow_field_affine": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "km_flow_field_tps": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "km_points_transform_affine": (_i, [_p, _p, _p, _i, _i, _p]),
    "km_points_transform_tps": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "km_warp_loss_workspace_bytes": (_sz, [_i, _i]),
    "km_warp_loss": (_i, [_i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "km_pair_stats_workspace_bytes": (_sz, [_i, _i, _ll, _i]),
    "km_pair_stats": (_i, [_p, _p, _p, _p, _i, _i, _ll, _i, _p]),
    "km_argmax_channels": (_i, [_p, _p, _i, _i, _ll, _p]),
    "km_com3d_workspace_bytes": (_sz, [_i, _i]),
    "km_com3d": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "km_fit_affine": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "km_fit_rigid": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "km_inverse44": (_i, [_p, _p, _p, _i, _p]),
    "km_tps_fit_workspace_bytes": (_sz, [_i, _i]),
    "km_tps_fit": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "km_pack_weights": (_i, [_p, _p, _i, _i, _i, _p]),
    "km_norm_finalize": (_i, [_p, _i, _i, _d, _p, _i, _i, _d, _d, _p, _p, _i, _f, _p, _p, _i, _p]),
    "km_channel_stats": (_i, [_p, _p, _i, _i, _ll, _p]),
    "km_norm_apply": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "km_pool_nparts": (_i, []),
    "km_maxpool2_stats": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "km_volume_stats": (_i, [_p, _p, _i, _ll, _p]),
    "km_stem_nparts": (_i, [_i, _i, _i, _i]),
    "km_conv3d_zfold_supported": (_i, [_i, _i, _i, _i, _i]),
    "km_pack_weights_zfold_bytes": (C.c_size_t, [_i, _i]),
    "km_pack_weights_zfold": (_i, [_p, _p, _i, _i, _p]),
    "km_conv3d_zfold": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "km_warp_labels_workspace_bytes": (C.c_size_t, [_i, _i]),
    "km_warp_labels_dice": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "km_jacobian_stats_workspace_bytes": (C.c_size_t, [_i]),
    "km_jacobian_stats": (_i, [_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, _p, _p, _i, _i, _i, _i, _p]),
    "km_conv3d_zfold_gn_workspace_bytes": (C.c_size_t, [_i]),
    "km_conv3d_zfold_gn": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "km_conv3d_zfold_pair_gn_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "km_conv3d_zfold_pair_gn": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "km_conv3d_tc_pair_gn_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "km_conv3d_tc_pair_gn": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "km_conv3d_zfold_pair_gn_cat": (_i, [_p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "km_conv3d_up2_supported": (_i, [_i, _i, _i, _i, _i]),
    "km_conv3d_up2_gn_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "km_conv3d_up2_gn": (_i, [_p, _p, _p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _p]),
    "km_conv3d_zfold_pair_gn_add": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "km_conv3d_tc_pair_gn_add": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "km_upsample2_ndhwc": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "km_hausdorff_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "km_hausdorff": (_i, [_p, _p, C.c_longlong, C.c_longlong, _i, _i, _i, _i, C.c_float, C.c_float, C.c_float, _p, _p, _p]),
    "km_conv3d_tc_pair_supported": (_i, [_i, _i, _i, _i, _i]),
    "km_conv3d_tc_pair": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "km_conv3d_zfold_pair_supported": (_i, [_i, _i, _i, _i, _i]),
    "km_pack_weights_zfold_pair": (_i, [_p, _p, _i, _i, _p]),
    "km_conv3d_zfold_pair": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "km_conv1x1_com_nparts": (_i, []),
    "km_conv1x1_com": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "km_conv3d_stem": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "km_conv_nparts": (_i, []),
    "km_conv3d_tc": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "km_com_finalize": (_i, [_p, _i, _p, _p, _i, _i, _p]),
    "km_ndhwc_bf16_to_ncdhw_f32": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "km_ncdhw_f32_to_ndhwc_bf16": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
}

KM_CONV_RELU, KM_CONV_STATS, KM_CONV_COM = 1, 2, 4
KM_INTERP_BILINEAR, KM_INTERP_NEAREST = 0, 1
KM_COORD_AFFINE, KM_COORD_TPS, KM_COORD_GRID = 0, 1, 2
KM_OPT_TPS_FAST = 1
KM_OPT_CONV_FORCE_GENERIC = 2
KM_OPT_CONV_NO_RESIDENT_WEIGHTS = 3
KM_OPT_CONV_MAX_BRICKS = 4
KM_OPT_CONV_NO_EPILOGUE_BATCH = 5
KM_OPT_CONV_HALO_AXIS = 6
KM_OPT_TPS_SINGLE_CTA = 7
KM_OPT_CONV_INTERLEAVE_BRICKS = 8
KM_OPT_CONV_TWO_ISSUERS = 9
KM_OPT_ZF2_TWO_BRICKS = 10
KM_OPT_TPS_PACKED = 11
KM_OPT_TPS_VPT = 12
KM_OPT_OPERAND_FP16 = 13
KM_OPT_WARP_TILE = 14

_lib = None


class KMError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libkm_b200.so (built by keymorph_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KMError(
            f"{LIB_PATH} is missing: run `python -m keymorph_b200.build` (there is no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


# kernels launched by one call of each entry point (bench.py's "gpu_launches" claim)
KERNELS_PER_CALL = {
    "km_warp_loss": 2, "km_pair_stats": 2, "km_com3d": 2, "km_tps_fit": 1, "km_warp_labels_dice": 3, "km_jacobian_stats": 2, "km_hausdorff": 13, "km_conv3d_zfold_gn": 2, "km_conv3d_zfold_pair_gn": 2, "km_conv3d_zfold_pair_gn_cat": 2, "km_conv3d_tc_pair_gn": 2, "km_conv3d_up2_gn": 2, "km_conv3d_zfold_pair_gn_add": 2, "km_conv3d_tc_pair_gn_add": 2,
}
launch_count = 0
# optional tracer: callable(name, phase) with phase in {"pre", "post"}; bench.py installs one that
# records CUDA events around km_conv3d_tc / km_warp_loss launches on the launching stream
TRACE = None


def call(name: str, *args):
    """Call an int-returning entry point and raise KMError with km_last_error() on failure."""
    global launch_count
    lib = load()
    if TRACE is not None:
        TRACE(name, "pre")
    rc = getattr(lib, name)(*args)
    if TRACE is not None:
        TRACE(name, "post")
    if rc != 0:
        raise KMError(f"{name} failed ({rc}): {lib.km_last_error().decode()}")
    launch_count += KERNELS_PER_CALL.get(name, 1)


def query(name: str, *args):
    return getattr(load(), name)(*args)
