"""ctypes face of oracle/_ref/libpointnet2_ref.so -- the reference's OWN lib/src CUDA kernels, compiled
unmodified for sm_100a by oracle/build_oracle.py --ref (TEST INFRASTRUCTURE; GPU box only).
Used (a) to pin the CUDA kernels and the C restatement bit-exactly against the real thing and (b) as the
"reference CUDA build" timing arm of individual operators.  Never imported by cmflow_b200/."""
import ctypes
import os

import torch

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libpointnet2_ref.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
    return _lib


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros(B, M, nsample, dtype=torch.int32, device=xyz.device)
    lib().ref_ball_query(B, N, M, ctypes.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _s())
    return idx


def group_points(points, idx):
    B, C, N = points.shape
    _, P, S = idx.shape
    out = torch.empty(B, C, P, S, device=points.device)
    lib().ref_group_points(B, C, N, P, S, _p(points), _p(idx), _p(out), _s())
    return out


def knn(k, unknown, known):
    B, N, _ = unknown.shape
    M = known.shape[1]
    d2 = torch.empty(B, N, k, device=unknown.device)
    idx = torch.empty(B, N, k, dtype=torch.int32, device=unknown.device)
    lib().ref_knn(B, N, M, k, _p(unknown), _p(known), _p(d2), _p(idx), _s())
    return d2, idx


def three_nn(unknown, known):
    B, N, _ = unknown.shape
    M = known.shape[1]
    d2 = torch.empty(B, N, 3, device=unknown.device)
    idx = torch.empty(B, N, 3, dtype=torch.int32, device=unknown.device)
    lib().ref_three_nn(B, N, M, _p(unknown), _p(known), _p(d2), _p(idx), _s())
    return d2, idx


def three_interpolate(points, idx, weight):
    B, C, M = points.shape
    N = idx.shape[1]
    out = torch.empty(B, C, N, device=points.device)
    lib().ref_three_interpolate(B, C, M, N, _p(points), _p(idx), _p(weight), _p(out), _s())
    return out


def gather_points(points, idx):
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty(B, C, M, device=points.device)
    lib().ref_gather_points(B, C, N, M, _p(points), _p(idx), _p(out), _s())
    return out


def furthest_point_sample(xyz, npoint):
    B, N, _ = xyz.shape
    temp = torch.full((B, N), 1e10, device=xyz.device)
    idx = torch.zeros(B, npoint, dtype=torch.int32, device=xyz.device)
    lib().ref_furthest_point_sampling(B, N, npoint, _p(xyz), _p(temp), _p(idx), _s())
    return idx, temp
