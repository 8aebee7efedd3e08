"""Drop-in for the reference's compiled extension module `pointnet2_cuda`.

Exports the ten functions of lib/src/pointnet2_api.cpp:11-24 with identical positional signatures and
the reference's ownership rule: the CALLER allocates every output tensor and passes it in; the callee
writes in place on the current CUDA stream and allocates nothing.  Put
`cmflow_b200/shim` on sys.path (it holds a one-line `pointnet2_cuda.py` re-exporting this module) and
the reference's lib/pointnet2_utils.py runs unmodified on these kernels.

Differences by design: a failing launch raises CmfError instead of `exit(-1)`
(lib/src/ball_query_gpu.cu:62-66); non-contiguous / non-CUDA / wrong-dtype tensors raise instead of
being read as garbage (the reference checks only in ball_query.cpp:16-17).
"""
import torch

from ._lib import check, dptr, lib, stream_ptr

_F, _I = torch.float32, torch.int32


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    check(lib().cmf_ball_query(b, n, m, float(radius), nsample, dptr(new_xyz, _F), dptr(xyz, _F), dptr(idx, _I), stream_ptr()))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    check(lib().cmf_group_points(b, c, n, npoints, nsample, dptr(points, _F), dptr(idx, _I), dptr(out, _F), stream_ptr()))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    check(lib().cmf_group_points_grad(b, c, n, npoints, nsample, dptr(grad_out, _F), dptr(idx, _I), dptr(grad_points, _F), stream_ptr()))
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    check(lib().cmf_gather_points(b, c, n, npoints, dptr(points, _F), dptr(idx, _I), dptr(out, _F), stream_ptr()))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    check(lib().cmf_gather_points_grad(b, c, n, npoints, dptr(grad_out, _F), dptr(idx, _I), dptr(grad_points, _F), stream_ptr()))
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    check(lib().cmf_furthest_point_sampling(b, n, m, dptr(points, _F), dptr(temp, _F), dptr(idx, _I), stream_ptr()))
    return 1


def knn_wrapper(b, n, m, k, unknown, known, dist2, idx):
    check(lib().cmf_knn(b, n, m, k, dptr(unknown, _F), dptr(known, _F), dptr(dist2, _F), dptr(idx, _I), stream_ptr()))


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    check(lib().cmf_three_nn(b, n, m, dptr(unknown, _F), dptr(known, _F), dptr(dist2, _F), dptr(idx, _I), stream_ptr()))


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    check(lib().cmf_three_interpolate(b, c, m, n, dptr(points, _F), dptr(idx, _I), dptr(weight, _F), dptr(out, _F), stream_ptr()))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    check(lib().cmf_three_interpolate_grad(b, c, n, m, dptr(grad_out, _F), dptr(idx, _I), dptr(weight, _F), dptr(grad_points, _F), stream_ptr()))


__all__ = ["ball_query_wrapper", "group_points_wrapper", "group_points_grad_wrapper", "gather_points_wrapper",
           "gather_points_grad_wrapper", "furthest_point_sampling_wrapper", "knn_wrapper", "three_nn_wrapper",
           "three_interpolate_wrapper", "three_interpolate_grad_wrapper"]
