"""ctypes binding of libnerfloc_b200.so (the C ABI declared in include/nerfloc_b200.h).

There is no CPU fallback: if the library cannot be loaded every product entry point raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NLB_LIB") or os.path.join(_HERE, "libnerfloc_b200.so")   # NLB_LIB: A/B builds of the same ABI

c_void_p, c_int, c_int64, c_size_t, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float
c_uint64 = ctypes.c_uint64


class NlbScene(ctypes.Structure):
    _fields_ = [("V", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("h", ctypes.c_int32),
                ("w", ctypes.c_int32), ("vh", ctypes.c_int32), ("vw", ctypes.c_int32), ("near_plane", c_float), ("far_plane", c_float),
                ("images", c_void_p), ("featmaps", c_void_p), ("vis_maps", c_void_p), ("cams", c_void_p),
                ("M", c_int64), ("sup_pre", c_void_p), ("sup_geo", c_void_p), ("knn_index", c_void_p),
                ("query_center", c_float * 3), ("featmaps_blend", c_void_p)]


# name -> (restype, argtypes); mirrors include/nerfloc_b200.h one to one
SIGNATURES = {
    "nlb_last_error": (ctypes.c_char_p, []),
    "nlb_version": (c_int, []),
    "nlb_knn_index_bytes": (c_size_t, [c_int64]),
    "nlb_knn_build": (c_int, [c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "nlb_knn_query": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "nlb_render_param_count": (c_int, []),
    "nlb_render_weights_floats": (c_size_t, [c_int]),
    "nlb_render_pack_weights": (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "nlb_support_prepare": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                    c_void_p, c_void_p]),
    "nlb_blend_prepare": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "nlb_query_scratch_bytes": (c_size_t, [c_int64, c_int]),
    "nlb_query_points": (c_int, [ctypes.POINTER(NlbScene), c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "nlb_aggregate_points": (c_int, [ctypes.POINTER(NlbScene), c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "nlb_descriptor_head": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "nlb_confidence_head": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "nlb_backproject_points": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nlb_render_scratch_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "nlb_render_rays": (c_int, [ctypes.POINTER(NlbScene), c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nlb_render_rays_gather": (c_int, [ctypes.POINTER(NlbScene), c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                       c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_size_t, c_void_p, c_int, c_int64, c_void_p]),
    "nlb_render_launch_count": (c_int64, [c_int64, c_int64]),
    "nlb_hierarchical_depths": (c_int, [ctypes.POINTER(NlbScene), c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                                        c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nlb_debug_read_prof": (c_int, [c_void_p, c_int]),
    "nlb_debug_knn_rays": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                                   c_void_p]),
    "nlb_debug_tc_gemm": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "nlb_profile_enable": (None, [c_int]),
    "nlb_profile_report": (c_int, [ctypes.c_char_p, c_size_t]),
    "nlb_match_weights_floats": (c_size_t, [c_int]),
    "nlb_match_pack_weights": (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "nlb_s2d_scores": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "nlb_mutual_scratch_bytes": (c_size_t, [c_int64, c_int64]),
    "nlb_mutual_matches": (c_int, [c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    "nlb_fine_windows": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64,
                                 c_void_p, c_void_p]),
    "nlb_fine_match": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "nlb_pnp_scratch_bytes": (c_size_t, [c_int]),
    "nlb_pnp_ransac": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_float, c_int, c_uint64, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


def load():
    """Returns the loaded library or raises (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m nerf_loc_b200.build` (needs nvcc); "
                               "nerf_loc_b200 has no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("nerfloc_b200: " + load().nlb_last_error().decode())


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("nerfloc_b200: expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("nerfloc_b200: tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def f32(t, device=None):
    t = t.detach()
    if device is not None:
        t = t.to(device)
    return t.to(torch.float32).contiguous()
