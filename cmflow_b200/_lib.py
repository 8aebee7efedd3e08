"""ctypes binding of libcmflow_b200.so (the C ABI of include/cmflow_b200.h).

No fallback: if the library is missing, `lib()` raises -- the product path must fail loudly rather
than silently compute somewhere else.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CMF_LIB: developer A/B switch -- another build of the SAME library (cmflow_b200.build.build_variant), never a different implementation
LIB_PATH = os.environ.get("CMF_LIB") or os.path.join(_HERE, "libcmflow_b200.so")
_lib = None

_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/cmflow_b200.h
SIGNATURES = {
    "cmf_last_error": [],
    "cmf_version": [],
    "cmf_device_check": [],
    "cmf_ball_query": [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp],
    "cmf_group_points": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "cmf_group_points_grad": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "cmf_gather_points": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "cmf_gather_points_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "cmf_furthest_point_sampling": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "cmf_knn": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "cmf_three_nn": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "cmf_three_interpolate": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "cmf_three_interpolate_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "cmf_knn_point": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "cmf_ball_query_ms": [_i, _i, _vp, _vp, _vp],
    "cmf_kde_density": [_i, _i, _i, _vp, _vp, _f, _vp, _vp],
    "cmf_kabsch_refine": [_i, _i, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp],
    "cmf_weighted_kabsch": [_i, _i, _vp, _vp, _vp, _vp, _vp],
    "cmf_model_blob_floats": [_i],
    "cmf_model_create": [ctypes.POINTER(_vp), _vp, _sz, _i, _f],
    "cmf_model_destroy": [_vp],
    "cmf_model_workspace_bytes": [_vp],
    "cmf_model_host_graphs": [_vp],
    "cmf_model_launches_per_forward": [_vp],
    "cmf_watchdog_read": [_vp],
    "cmf_model_forward": [_vp, _i, _i] + [_vp] * 11,
    "cmf_model_forward_host": [_vp, _i, _i] + [_vp] * 11,
    "cmf_model_forward2": [_vp, _i, _i, _i] + [_vp] * 12,
    "cmf_model_forward_raflow2": [_vp, _i, _i, _i] + [_vp] * 10,
    "cmf_model_forward_host2": [_vp, _i, _i, _i] + [_vp] * 11,
    "cmf_model_submit_host": [_vp, _i, _i, _i, _i] + [_vp] * 11,
    "cmf_model_wait_host": [_vp, _i],
    "cmf_model_tap": [_vp, ctypes.c_char_p],
    "cmf_eval_scene_flow_sums": [_i, _i, _vp, _vp, _vp, _vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, _vp, _vp],
    "cmf_eval_motion_seg_counts": [ctypes.c_longlong, _vp, _vp, _vp, _vp],
    "cmf_eval_rpe_sums": [_i, _vp, _vp, _vp, _vp],
    "cmf_model_set_raflow": [_vp, _f, _f],
    "cmf_model_forward_raflow": [_vp, _i, _i] + [_vp] * 10,
    "cmf_raflow_refine": [_i, _i, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp],
    "cmf_model_set_profiling": [_vp, _i],
    "cmf_model_set_mode": [_vp, _i],
    "cmf_model_get_mode": [_vp],
    "cmf_test_tc_gemm": [_i, _i, ctypes.c_longlong, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp],
    "cmf_test_tc_gemm_fmt": [_i, _i, _i, ctypes.c_longlong, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp],
    "cmf_test_tc_tiled_floats": [_i, _i],
    "cmf_test_tc_set_dbg": [_vp],
    "cmf_model_profile_categories": [],
    "cmf_model_profile_name": [_i],
    "cmf_model_read_profile": [_vp, _vp, _vp, _vp],
}
_RESTYPES = {"cmf_last_error": ctypes.c_char_p, "cmf_version": ctypes.c_char_p, "cmf_model_blob_floats": _sz,
             "cmf_model_destroy": None, "cmf_test_tc_set_dbg": None, "cmf_model_profile_name": ctypes.c_char_p, "cmf_test_tc_tiled_floats": _sz, "cmf_model_workspace_bytes": _sz, "cmf_model_tap": _vp}


class CmfError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CmfError(f"{LIB_PATH} is not built -- run `python -m cmflow_b200.build` "
                           "(nvcc, sm_100a). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)           # AttributeError if the header and the .so ever disagree
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _lib = L
    return _lib


def watchdog_record():
    """None, or where a tensor-core kernel's mbarrier watchdog fired (see cmf_watchdog_read): readable after the launch failure."""
    out = (ctypes.c_ulonglong * 4)()
    lib().cmf_watchdog_read(out)
    if not out[0]:
        return None
    fam = {1: "tc_gemm (one-CTA GEMM)", 2: "tc_gemm2 (CTA-pair GEMM)", 3: "tc_sc2 (fused set-conv #2)", 4: "tc_chain (set-conv #1 / mlp2)"}.get(out[0], str(out[0]))
    return {"kernel_family": fam, "block": out[1] >> 32, "thread": out[1] & 0xffffffff, "warp": (out[1] & 0xffffffff) >> 5,
            "grid": out[2] >> 32, "block_size": out[2] & 0xffffffff, "waited_s": out[3] / 1e9}


def check(rc):
    if rc != 0:
        msg = lib().cmf_last_error().decode()
        wd = watchdog_record()
        if wd:
            msg += f" [mbarrier watchdog fired: {wd}]"
        raise CmfError(msg)


def stream_ptr():
    """The current torch CUDA stream as a cudaStream_t (the reference enqueues on
    at::cuda::getCurrentCUDAStream(), e.g. lib/src/ball_query.cpp:22)."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise CmfError("expected a CUDA tensor (this library has no CPU path)")
    if not t.is_contiguous():
        raise CmfError("expected a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise CmfError(f"expected dtype {dtype}, got {t.dtype}")
    return ctypes.c_void_p(t.data_ptr())
