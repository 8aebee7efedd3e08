"""Operator API of the reference's lib/pointnet2_utils.py on the B200 kernels.

Same callables, argument meaning, shapes, dtypes and autograd behaviour:
  furthest_point_sample, gather_operation, knn, three_nn, three_interpolate, grouping_operation,
  ball_query, QueryAndGroup, GroupAll            (lib/pointnet2_utils.py:10-318)
Inputs must be contiguous CUDA tensors (the reference asserts contiguity, e.g. :22,50-51); index
tensors are int32.  Everything runs on the current CUDA stream through the C ABI; there is no CPU path.
"""
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from . import pointnet2_cuda as _k


def _new(like, *shape, dtype=torch.float32, zero=False):
    f = torch.zeros if zero else torch.empty
    return f(*shape, dtype=dtype, device=like.device)


class _FurthestPointSampling(Function):          # lib/pointnet2_utils.py:10-37
    @staticmethod
    def forward(ctx, xyz, npoint):
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        out = _new(xyz, B, npoint, dtype=torch.int32)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        _k.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = _FurthestPointSampling.apply


class _GatherOperation(Function):                # lib/pointnet2_utils.py:40-72
    @staticmethod
    def forward(ctx, features, idx):
        assert features.is_contiguous() and idx.is_contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        out = _new(features, B, C, npoint)
        _k.gather_points_wrapper(B, C, N, npoint, features, idx, out)
        ctx.for_backwards = (idx, C, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad = _new(grad_out, B, C, N, zero=True)
        _k.gather_points_grad_wrapper(B, C, N, npoint, grad_out.contiguous(), idx, grad)
        return grad, None


gather_operation = _GatherOperation.apply


class _KNN(Function):                            # lib/pointnet2_utils.py:74-102
    @staticmethod
    def forward(ctx, k, unknown, known):
        assert unknown.is_contiguous() and known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = _new(unknown, B, N, k)
        idx = _new(unknown, B, N, k, dtype=torch.int32)
        _k.knn_wrapper(B, N, m, k, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None


knn = _KNN.apply


class _ThreeNN(Function):                        # lib/pointnet2_utils.py:104-135
    @staticmethod
    def forward(ctx, unknown, known):
        assert unknown.is_contiguous() and known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = _new(unknown, B, N, 3)
        idx = _new(unknown, B, N, 3, dtype=torch.int32)
        _k.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = _ThreeNN.apply


class _ThreeInterpolate(Function):               # lib/pointnet2_utils.py:138-184
    @staticmethod
    def forward(ctx, features, idx, weight):
        assert features.is_contiguous() and idx.is_contiguous() and weight.is_contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        out = _new(features, B, c, n)
        _k.three_interpolate_wrapper(B, c, m, n, features, idx, weight, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad = _new(grad_out, B, c, m, zero=True)
        _k.three_interpolate_grad_wrapper(B, c, n, m, grad_out.contiguous(), idx, weight, grad)
        return grad, None, None


three_interpolate = _ThreeInterpolate.apply


class _GroupingOperation(Function):              # lib/pointnet2_utils.py:187-225
    @staticmethod
    def forward(ctx, features, idx):
        assert features.is_contiguous() and idx.is_contiguous()
        idx = idx.int()
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        out = _new(features, B, C, nfeatures, nsample)
        _k.group_points_wrapper(B, C, N, nfeatures, nsample, features, idx, out)
        ctx.for_backwards = (idx, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad = _new(grad_out, B, C, N, zero=True)
        _k.group_points_grad_wrapper(B, C, N, npoint, nsample, grad_out.contiguous(), idx, grad)
        return grad, None


grouping_operation = _GroupingOperation.apply


class _BallQuery(Function):                      # lib/pointnet2_utils.py:228-255
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        assert new_xyz.is_contiguous() and xyz.is_contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = _new(xyz, B, npoint, nsample, dtype=torch.int32, zero=True)
        _k.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = _BallQuery.apply


class QueryAndGroup(nn.Module):                  # lib/pointnet2_utils.py:258-292
    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None) -> Tuple[torch.Tensor]:
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features


class GroupAll(nn.Module):                       # lib/pointnet2_utils.py:295-318
    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
