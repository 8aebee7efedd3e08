"""ctypes face of oracle/pointops_oracle.c (TEST INFRASTRUCTURE) over CPU torch tensors.

Also provides `as_pointnet2_module()`: a module object with the ten functions the reference's
lib/pointnet2_utils.py expects from `pointnet2_cuda` (lib/src/pointnet2_api.cpp:11-24), backed
by the C restatement, so the *unmodified* reference Python can run on CPU when generating golden
vectors (tests/golden/make_golden.py).
"""
import ctypes
import types

import torch

from .build_oracle import build_c

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
    return _lib


def _p(t):
    assert t.device.type == "cpu" and t.is_contiguous(), "oracle works on contiguous CPU tensors"
    return ctypes.c_void_p(t.data_ptr())


def _f32(t):
    assert t.dtype == torch.float32
    return _p(t)


def _i32(t):
    assert t.dtype == torch.int32
    return _p(t)


def ball_query(radius, nsample, xyz, new_xyz):
    """xyz (B,N,3), new_xyz (B,M,3) -> idx (B,M,nsample) int32 (zero-initialised like pointnet2_utils.py:246)."""
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros(B, M, nsample, dtype=torch.int32)
    lib().orc_ball_query(B, N, M, ctypes.c_float(radius), nsample, _f32(new_xyz), _f32(xyz), _i32(idx))
    return idx


def group_points(points, idx):
    B, C, N = points.shape
    _, P, S = idx.shape
    out = torch.empty(B, C, P, S, dtype=torch.float32)
    lib().orc_group_points(B, C, N, P, S, _f32(points), _i32(idx), _f32(out))
    return out


def group_points_grad(grad_out, idx, N):
    B, C, P, S = grad_out.shape
    gp = torch.zeros(B, C, N, dtype=torch.float32)
    lib().orc_group_points_grad(B, C, N, P, S, _f32(grad_out), _i32(idx), _f32(gp))
    return gp


def gather_points(points, idx):
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty(B, C, M, dtype=torch.float32)
    lib().orc_gather_points(B, C, N, M, _f32(points), _i32(idx), _f32(out))
    return out


def gather_points_grad(grad_out, idx, N):
    B, C, M = grad_out.shape
    gp = torch.zeros(B, C, N, dtype=torch.float32)
    lib().orc_gather_points_grad(B, C, N, M, _f32(grad_out), _i32(idx), _f32(gp))
    return gp


def knn(k, unknown, known):
    """lib kNN (interpolate_gpu.cu:9-57): returns (dist2 (B,N,k) f32, idx (B,N,k) i32), ascending."""
    B, N, _ = unknown.shape
    M = known.shape[1]
    d2 = torch.empty(B, N, k, dtype=torch.float32)
    idx = torch.empty(B, N, k, dtype=torch.int32)
    rc = lib().orc_knn(B, N, M, k, _f32(unknown), _f32(known), _f32(d2), _i32(idx))
    if rc:
        raise ValueError("k must be <= 200")
    return d2, idx


def three_nn(unknown, known):
    B, N, _ = unknown.shape
    M = known.shape[1]
    d2 = torch.empty(B, N, 3, dtype=torch.float32)
    idx = torch.empty(B, N, 3, dtype=torch.int32)
    lib().orc_three_nn(B, N, M, _f32(unknown), _f32(known), _f32(d2), _i32(idx))
    return d2, idx


def three_interpolate(points, idx, weight):
    B, C, M = points.shape
    N = idx.shape[1]
    out = torch.empty(B, C, N, dtype=torch.float32)
    lib().orc_three_interpolate(B, C, M, N, _f32(points), _i32(idx), _f32(weight), _f32(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, M):
    B, C, N = grad_out.shape
    gp = torch.zeros(B, C, M, dtype=torch.float32)
    lib().orc_three_interpolate_grad(B, C, N, M, _f32(grad_out), _i32(idx), _f32(weight), _f32(gp))
    return gp


def furthest_point_sample(xyz, npoint):
    B, N, _ = xyz.shape
    temp = torch.full((B, N), 1e10, dtype=torch.float32)
    idx = torch.zeros(B, npoint, dtype=torch.int32)
    lib().orc_furthest_point_sampling(B, N, npoint, _f32(xyz), _f32(temp), _i32(idx))
    return idx


def knn_point(nsample, xyz, new_xyz):
    """Model kNN (radarflow_util.py:88-99): xyz (B,N,3) candidates, new_xyz (B,S,3) queries ->
    (idx (B,S,k) int32 ascending by (d, index), dist (B,S,k) f32 in the expanded form)."""
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = torch.empty(B, S, nsample, dtype=torch.int32)
    dist = torch.empty(B, S, nsample, dtype=torch.float32)
    rc = lib().orc_knn_point(B, N, S, nsample, _f32(xyz), _f32(new_xyz), _i32(idx), _f32(dist))
    if rc:
        raise ValueError("need 1 <= k <= min(N, 64)")
    return idx, dist


def as_pointnet2_module():
    """A stand-in for the compiled `pointnet2_cuda` extension with the reference's exact positional
    signatures (pointnet2_api.cpp:11-24), for running the reference Python on CPU."""
    m = types.ModuleType("pointnet2_cuda")
    L = lib()

    def ball_query_wrapper(b, n, m_, radius, nsample, new_xyz, xyz, idx):
        L.orc_ball_query(b, n, m_, ctypes.c_float(radius), nsample, _f32(new_xyz), _f32(xyz), _i32(idx))
        return 1

    def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
        L.orc_group_points(b, c, n, npoints, nsample, _f32(points), _i32(idx), _f32(out))
        return 1

    def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
        L.orc_group_points_grad(b, c, n, npoints, nsample, _f32(grad_out), _i32(idx), _f32(grad_points))
        return 1

    def gather_points_wrapper(b, c, n, npoints, points, idx, out):
        L.orc_gather_points(b, c, n, npoints, _f32(points), _i32(idx), _f32(out))
        return 1

    def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
        L.orc_gather_points_grad(b, c, n, npoints, _f32(grad_out), _i32(idx), _f32(grad_points))
        return 1

    def furthest_point_sampling_wrapper(b, n, m_, points, temp, idx):
        L.orc_furthest_point_sampling(b, n, m_, _f32(points), _f32(temp), _i32(idx))
        return 1

    def knn_wrapper(b, n, m_, k, unknown, known, dist2, idx):
        L.orc_knn(b, n, m_, k, _f32(unknown), _f32(known), _f32(dist2), _i32(idx))

    def three_nn_wrapper(b, n, m_, unknown, known, dist2, idx):
        L.orc_three_nn(b, n, m_, _f32(unknown), _f32(known), _f32(dist2), _i32(idx))

    def three_interpolate_wrapper(b, c, m_, n, points, idx, weight, out):
        L.orc_three_interpolate(b, c, m_, n, _f32(points), _i32(idx), _f32(weight), _f32(out))

    def three_interpolate_grad_wrapper(b, c, n, m_, grad_out, idx, weight, grad_points):
        L.orc_three_interpolate_grad(b, c, n, m_, _f32(grad_out), _i32(idx), _f32(weight), _f32(grad_points))

    for f in (ball_query_wrapper, group_points_wrapper, group_points_grad_wrapper, gather_points_wrapper,
              gather_points_grad_wrapper, furthest_point_sampling_wrapper, knn_wrapper, three_nn_wrapper,
              three_interpolate_wrapper, three_interpolate_grad_wrapper):
        setattr(m, f.__name__, f)
    return m
