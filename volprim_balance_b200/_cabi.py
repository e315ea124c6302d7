"""ctypes binding of libvolprim_cuda.so (include/volprim_cuda.h).

The library is the ONLY compute path of this package: there is no CPU or PyTorch fallback.  Importing the
package works without a GPU (so that host-side logic can be tested), but every call that needs the
kernels raises `VolprimCudaError` if the shared library is missing or the device is not a B200.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VOLPRIM_CUDA_LIB", os.path.join(_PKG, "libvolprim_cuda.so"))  # override: kernel tuning experiments

VP_OK = 0
INTEGRATOR_RF, INTEGRATOR_TOMO = 0, 1
KERNEL_GAUSSIAN, KERNEL_EPANECHNIKOV = 0, 1
MAX_DEPTH_UNLIMITED = 0xFFFFFFFF


class VolprimCudaError(RuntimeError):
    pass


class vp_params(C.Structure):
    _fields_ = [
        ("integrator", C.c_int32),
        ("kernel", C.c_int32),
        ("max_depth", C.c_uint32),
        ("srgb_primitives", C.c_int32),
        ("hide_emitters", C.c_int32),
        ("t_cutoff", C.c_float),
        ("eps_advance", C.c_float),
        ("env", C.c_float * 3),
        ("image_width", C.c_int32),
        ("image_height", C.c_int32),
        ("use_rr", C.c_int32),
        ("rr_depth", C.c_uint32),
        ("rr_seed", C.c_uint32),
        ("rr_skip", C.c_uint32),
    ]


class vp_camera(C.Structure):
    _fields_ = [
        ("to_world", C.c_float * 12),
        ("fov_x_deg", C.c_float),
        ("near_clip", C.c_float),
        ("far_clip", C.c_float),
        ("cx", C.c_float),
        ("cy", C.c_float),
        ("width", C.c_int32),
        ("height", C.c_int32),
    ]


class vp_ray_source(C.Structure):
    _fields_ = [
        ("ray_o", C.c_void_p),
        ("ray_d", C.c_void_p),
        ("ray_maxt", C.c_void_p),
        ("camera", C.POINTER(vp_camera)),
        ("jitter", C.c_void_p),
        ("spp", C.c_int32),
        ("row_begin", C.c_int32),
        ("row_count", C.c_int32),
        ("reserved", C.c_int32),
    ]


class vp_hit_record(C.Structure):
    _fields_ = [
        ("ray_offsets", C.c_void_p),
        ("ids", C.c_void_p),
        ("state", C.c_void_p),
        ("counts", C.c_void_p),
        ("total", C.c_void_p),
        ("capacity", C.c_int64),
        ("id_cap", C.c_int32),
        ("dense", C.c_int32),
    ]


RFILTER_BOX, RFILTER_TENT, RFILTER_GAUSSIAN = 0, 1, 2
RFILTERS = {"box": RFILTER_BOX, "tent": RFILTER_TENT, "gaussian": RFILTER_GAUSSIAN}


class vp_stats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64),
        ("hits", C.c_uint64),
        ("candidates", C.c_uint64),
        ("node_visits", C.c_uint64),
        ("passes", C.c_uint64),
        ("stack_overflows", C.c_uint64),
        ("interval_retries", C.c_uint64),
    ]


# name -> (restype, argtypes); every symbol include/volprim_cuda.h declares
_VP = C.c_void_p
SIGNATURES = {
    "vp_version": (C.c_int, []),
    "vp_create": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "vp_destroy": (C.c_int, [_VP]),
    "vp_last_error": (C.c_char_p, [_VP]),
    "vp_set_primitives": (C.c_int, [_VP, C.c_int64, _VP, _VP, _VP, C.c_int32, C.c_float, _VP]),
    "vp_build": (C.c_int, [_VP, _VP]),
    "vp_refit": (C.c_int, [_VP, _VP]),
    "vp_trace_forward": (C.c_int, [_VP, C.POINTER(vp_params), C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP, _VP,
                                   C.c_int32, C.c_int64, C.c_int64, _VP]),
    "vp_trace_adjoint": (C.c_int, [_VP, C.POINTER(vp_params), C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP, _VP,
                                   C.c_int32, C.c_int64, C.c_int64, _VP, _VP, _VP, _VP]),
    "vp_raygen_perspective": (C.c_int, [_VP, C.POINTER(vp_camera), C.c_int32, _VP, _VP, _VP, _VP, _VP]),
    "vp_render_forward": (C.c_int, [_VP, C.POINTER(vp_params), C.POINTER(vp_ray_source), C.c_int64, _VP, _VP, _VP,
                                    C.POINTER(vp_hit_record), _VP]),
    "vp_render_adjoint": (C.c_int, [_VP, C.POINTER(vp_params), C.POINTER(vp_ray_source), C.c_int64, _VP, _VP,
                                    C.POINTER(vp_hit_record), _VP, _VP, _VP, _VP]),
    "vp_adjoint_begin": (C.c_int, [_VP, C.POINTER(vp_params), C.POINTER(vp_ray_source), C.c_int64, _VP, _VP,
                                   C.POINTER(vp_hit_record), _VP, _VP, _VP, _VP]),
    "vp_adjoint_finish": (C.c_int, [_VP, C.POINTER(vp_params), C.POINTER(vp_ray_source), C.c_int64,
                                    C.POINTER(vp_hit_record), C.c_int64, C.c_int64, _VP, _VP, _VP, _VP]),
    "vp_film_splat": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP]),
    "vp_film_develop": (C.c_int, [C.c_int32, C.c_int32, _VP, _VP, C.c_int64, _VP]),
    "vp_film_adjoint": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP, C.c_int64, _VP, _VP]),
    "vp_set_option": (C.c_int, [_VP, C.c_char_p, C.c_int64]),
    "vp_get_stats": (C.c_int, [_VP, C.POINTER(vp_stats), _VP]),
    "vp_bounded_adam_step": (C.c_int, [C.c_int64, _VP, _VP, _VP, _VP, C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.c_int, C.c_float, C.c_int, C.c_float, _VP]),
    "vp_l1_loss_grad": (C.c_int, [C.c_int64, _VP, _VP, C.c_double, _VP, _VP, _VP]),
    "vp_debug_bvh": (C.c_int, [_VP, _VP, _VP, C.POINTER(C.c_int64), _VP]),
    "vp_debug_selftest": (C.c_int, [C.c_int32, C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]),
}

_lib = None


def load_library():
    """Load libvolprim_cuda.so (built by `make -C volprim_balance_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VolprimCudaError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the volprim integrators)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, ctx=None) -> None:
    if rc != VP_OK:
        msg = load_library().vp_last_error(ctx)
        raise VolprimCudaError(f"libvolprim_cuda error {rc}: {msg.decode() if msg else '?'}")
