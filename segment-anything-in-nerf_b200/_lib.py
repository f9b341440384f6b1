"""ctypes binding of ``libsnrf.so`` (C ABI declared in ``include/snrf.h``).

The CUDA library is the product: there is no Python / PyTorch fallback.  Importing this module without the
built library, or creating a context without a B200, raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SNRF_LIB_PATH") or os.path.join(_HERE, "libsnrf.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["api.cu", "march.cu", "march_v1.cu", "sam.cu", "gemm.cu", "query.cu", "raygen.cu", "backward.cu", "sam_bucket.cu", "bricks.cu", "gemm_tma.cu", "exchange.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--threads", "0",  # one compile job per source file
]

SNRF_MAX_LEVELS = 16
WANT_SAM, WANT_CLIPSEG, PATCH = 1, 2, 4
BG_LAST_SAMPLE, BG_FIXED = 0, 1


class Level(C.Structure):
    _fields_ = [("scale", C.c_float), ("res", C.c_uint32), ("size", C.c_uint32), ("offset", C.c_uint32),
                ("hashed", C.c_uint32)]


class GridDesc(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("n_features", C.c_int32), ("lv", Level * SNRF_MAX_LEVELS)]


class RenderOpts(C.Structure):
    _fields_ = [("near_plane", C.c_float), ("far_plane", C.c_float), ("hist_padding", C.c_float),
                ("bg_mode", C.c_int32), ("bg", C.c_float * 3), ("k_sam", C.c_int32), ("sharpen", C.c_float),
                ("patch_size", C.c_int32)]


class DebugOut(C.Structure):
    _fields_ = [("prop_weights", C.c_void_p), ("edges", C.c_void_p), ("weights", C.c_void_p),
                ("density", C.c_void_p), ("rgb_samples", C.c_void_p), ("sam_t", C.c_void_p), ("sam_w", C.c_void_p),
                ("sam_feat", C.c_void_p)]


class Camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("width", C.c_int32),
                ("height", C.c_int32), ("camera_type", C.c_int32), ("has_distortion", C.c_int32),
                ("distortion", C.c_float * 6), ("c2w", C.c_float * 12), ("has_aabb", C.c_int32), ("aabb", C.c_float * 6)]


# every symbol include/snrf.h declares: name -> (restype, argtypes)
_P, _I, _L, _U, _F = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_float
SYMBOLS = {
    "snrf_grid_desc_init": (_I, [C.POINTER(GridDesc), _I, _I, _I, _I, _F]),
    "snrf_ctx_create": (_I, [_I, C.POINTER(_P)]),
    "snrf_ctx_destroy": (None, [_P]),
    "snrf_last_error": (C.c_char_p, [_P]),
    "snrf_set_engine": (_I, [_P, _I]),
    "snrf_set_pdf_u": (_I, [_P, _P, _I]),
    "snrf_set_early_termination": (_I, [_P, _F]),
    "snrf_set_jitter": (_I, [_P, _P, _L]),
    "snrf_set_feature_cutoff": (_I, [_P, _F]),
    "snrf_set_brick_budget": (_I, [_P, _L, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "snrf_set_anneal": (_I, [_P, _F]),
    "snrf_feature_slot_stats": (_I, [_P, C.POINTER(C.c_int64), _I]),
    "snrf_upload_proposal": (_I, [_P, _P, _L, C.POINTER(GridDesc), _P]),
    "snrf_upload_field_base": (_I, [_P, _P, _L, C.POINTER(GridDesc), _P]),
    "snrf_upload_field_head": (_I, [_P, _P, _L, _P]),
    "snrf_upload_feature_grid": (_I, [_P, _I, _I, _P, _L, C.POINTER(GridDesc), _P]),
    "snrf_upload_feature_net": (_I, [_P, _I, _P, _L, _I, _P]),
    "snrf_upload_conv_head": (_I, [_P, _P, _P, _P, _P, _P]),
    "snrf_render": (_I, [_P, _P, _P, _P, _P, _L, _U, C.POINTER(RenderOpts), _P, _P, _P, _P, _P, _P,
                         C.POINTER(DebugOut), _P]),
    "snrf_sample": (_I, [_P, _P, _P, _P, _P, _L, C.POINTER(RenderOpts), _P, _P, _P, _P]),
    "snrf_patch_aggregate": (_I, [_P, _P, _L, _I, _P, _P]),
    "snrf_query_density": (_I, [_P, _I, _P, _L, _P, _P, _P]),
    "snrf_query_rgb": (_I, [_P, _P, _P, _L, _P, _P]),
    "snrf_query_features": (_I, [_P, _I, _P, _L, _P, _P, _P]),
    "snrf_ray_op": (_I, [_P, _I, _P, _P, _P, _P, _L, _I, _I, _I, _P, _P]),
    "snrf_render_frame": (_I, [_P, _P, _P, _P, _P, _L, _L, _U, C.POINTER(RenderOpts), _P, _P, _P, _P, _P, _P, _P]),
    "snrf_set_pipeline": (_I, [_P, _I]),
    "snrf_set_replication": (_I, [_P, _I, _P, _L, _P, C.POINTER(C.c_void_p), _I]),
    "snrf_set_replication_mode": (_I, [_P, _I]),
    "snrf_set_feature_dtype": (_I, [_P, _I]),
    "snrf_set_march_first": (_I, [_P, _I]),
    "snrf_generate_rays": (_I, [_P, C.POINTER(Camera), _P, _I, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "snrf_render_camera": (_I, [_P, C.POINTER(Camera), _P, _I, _P, _I, _L, _U, C.POINTER(RenderOpts), _P, _P, _P, _P,
                                _P, _P, _P]),
    "snrf_feature_forward": (_I, [_P, _I, _P, _P, _P, _P, _L, _P, _P, _P]),
    "snrf_feature_backward": (_I, [_P, _I, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P]),
    "snrf_patch_aggregate_backward": (_I, [_P, _P, _L, _I, _P, _P, _P, _P, _P, _P, _P]),
    "snrf_field_backward": (_I, [_P, _I, _P, _P, _L, _P, _P, _P, _P, _P]),
    "snrf_pick_samples": (_I, [_P, _P, _P, _P, _L, _I, _I, _F, _P, _P, _P]),
    "snrf_ray_op_backward": (_I, [_P, _I, _P, _P, _P, _P, _P, _L, _I, _I, _P, _P]),
    "snrf_launch_count": (_L, [_P]),
    "snrf_set_timing": (_I, [_P, _I]),
    "snrf_kernel_times": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile ``csrc/*.cu`` for sm_100a into ``libsnrf.so`` next to this file (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in ("common.cuh", "kernels.cuh", "raygen.cuh", "backward.cuh")] + [
        os.path.join(_HERE, "..", "include", "snrf.h")
    ]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("SNRF_NVCC_EXTRA", "").split()  # e.g. -DSNRF_MARCH_MIN_CTAS=4 for a tuning variant
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load the library and bind every declared symbol.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no fallback path: the CUDA library is the product)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
