"""Operator API of the reference's lib/pointnet2_utils.py on the B200 kernels.

Same callables, argument meaning, shapes, dtypes and autograd behaviour:
  furthest_point_sample, gather_operation, knn, three_nn, three_interpolate, grouping_operation,
  ball_query, QueryAndGroup, GroupAll            (lib/pointnet2_utils.py:10-318)
Inputs must be contiguous CUDA tensors (the reference asserts contiguity, e.g. :22,50-51); index
tensors are int32.  Everything runs on the current CUDA stream through the C ABI; there is no CPU path.

Layout of this file: the seven operators differ only in (a) which launcher they call, (b) the shapes of
the outputs the CALLER allocates (the reference's ownership rule), and (c) whether a gradient flows
back to the first tensor argument -- so each is one small autograd.Function built on two helpers.
"""
import torch
from torch import nn
from torch.autograd import Function

from . import pointnet2_cuda as _k


def _out(ref, shape, dtype=torch.float32, fill=None):
    """Caller-allocated output next to `ref` (same device); `fill` = None leaves it uninitialised."""
    if fill is None:
        return torch.empty(shape, dtype=dtype, device=ref.device)
    return torch.full(shape, fill, dtype=dtype, device=ref.device)


def _contig(*tensors):
    for t in tensors:
        assert t.is_contiguous()


class _IndexOp(Function):
    """Base of the operators that produce indices (and distances): nothing to differentiate."""

    @staticmethod
    def backward(ctx, *grads):
        return (None,) * ctx.n_inputs


class _FurthestPointSampling(_IndexOp):          # lib/pointnet2_utils.py:10-37
    @staticmethod
    def forward(ctx, xyz, npoint):
        ctx.n_inputs = 2
        _contig(xyz)
        nb, npts = xyz.shape[0], xyz.shape[1]
        picked = _out(xyz, (nb, npoint), torch.int32)
        running_min = _out(xyz, (nb, npts), fill=1e10)
        _k.furthest_point_sampling_wrapper(nb, npts, npoint, xyz, running_min, picked)
        ctx.mark_non_differentiable(picked)
        return picked


class _KNN(_IndexOp):                            # lib/pointnet2_utils.py:74-102
    @staticmethod
    def forward(ctx, k, unknown, known):
        ctx.n_inputs = 3
        _contig(unknown, known)
        nb, nq, nc = unknown.shape[0], unknown.shape[1], known.shape[1]
        d2 = _out(unknown, (nb, nq, k))
        nbr = _out(unknown, (nb, nq, k), torch.int32)
        _k.knn_wrapper(nb, nq, nc, k, unknown, known, d2, nbr)
        ctx.mark_non_differentiable(nbr)
        return d2.sqrt(), nbr                      # the reference returns distances, not squares (:97)


class _ThreeNN(_IndexOp):                        # lib/pointnet2_utils.py:104-135
    @staticmethod
    def forward(ctx, unknown, known):
        ctx.n_inputs = 2
        _contig(unknown, known)
        nb, nq, nc = unknown.shape[0], unknown.shape[1], known.shape[1]
        d2 = _out(unknown, (nb, nq, 3))
        nbr = _out(unknown, (nb, nq, 3), torch.int32)
        _k.three_nn_wrapper(nb, nq, nc, unknown, known, d2, nbr)
        ctx.mark_non_differentiable(nbr)
        return d2.sqrt(), nbr


class _BallQuery(_IndexOp):                      # lib/pointnet2_utils.py:228-255
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        ctx.n_inputs = 4
        _contig(new_xyz, xyz)
        nb, npts, nq = xyz.shape[0], xyz.shape[1], new_xyz.shape[1]
        nbr = _out(xyz, (nb, nq, nsample), torch.int32, fill=0)      # pre-zeroed by the caller, as the kernel expects (:246)
        _k.ball_query_wrapper(nb, npts, nq, radius, nsample, new_xyz, xyz, nbr)
        ctx.mark_non_differentiable(nbr)
        return nbr


class _GatherOperation(Function):                # lib/pointnet2_utils.py:40-72
    @staticmethod
    def forward(ctx, features, idx):
        _contig(features, idx)
        nb, nch, npts = features.shape
        npick = idx.shape[1]
        picked = _out(features, (nb, nch, npick))
        _k.gather_points_wrapper(nb, nch, npts, npick, features, idx, picked)
        ctx.saved = (idx, nch, npts)
        return picked

    @staticmethod
    def backward(ctx, grad_out):
        idx, nch, npts = ctx.saved
        nb, npick = idx.shape
        g = _out(grad_out, (nb, nch, npts), fill=0.0)
        _k.gather_points_grad_wrapper(nb, nch, npts, npick, grad_out.contiguous(), idx, g)
        return g, None


class _ThreeInterpolate(Function):               # lib/pointnet2_utils.py:138-184
    @staticmethod
    def forward(ctx, features, idx, weight):
        _contig(features, idx, weight)
        nb, nch, nsrc = features.shape
        ndst = idx.shape[1]
        mixed = _out(features, (nb, nch, ndst))
        _k.three_interpolate_wrapper(nb, nch, nsrc, ndst, features, idx, weight, mixed)
        ctx.saved = (idx, weight, nsrc)
        return mixed

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, nsrc = ctx.saved
        nb, nch, ndst = grad_out.shape
        g = _out(grad_out, (nb, nch, nsrc), fill=0.0)
        _k.three_interpolate_grad_wrapper(nb, nch, ndst, nsrc, grad_out.contiguous(), idx, weight, g)
        return g, None, None


class _GroupingOperation(Function):              # lib/pointnet2_utils.py:187-225
    @staticmethod
    def forward(ctx, features, idx):
        _contig(features, idx)
        idx = idx.int()
        nb, nq, ns = idx.shape
        nch, npts = features.shape[1], features.shape[2]
        grouped = _out(features, (nb, nch, nq, ns))
        _k.group_points_wrapper(nb, nch, npts, nq, ns, features, idx, grouped)
        ctx.saved = (idx, npts)
        return grouped

    @staticmethod
    def backward(ctx, grad_out):
        idx, npts = ctx.saved
        nb, nch, nq, ns = grad_out.shape
        g = _out(grad_out, (nb, nch, npts), fill=0.0)
        _k.group_points_grad_wrapper(nb, nch, npts, nq, ns, grad_out.contiguous(), idx, g)
        return g, None


furthest_point_sample = _FurthestPointSampling.apply
gather_operation = _GatherOperation.apply
knn = _KNN.apply
three_nn = _ThreeNN.apply
three_interpolate = _ThreeInterpolate.apply
grouping_operation = _GroupingOperation.apply
ball_query = _BallQuery.apply


class QueryAndGroup(nn.Module):                  # lib/pointnet2_utils.py:258-292
    """Ball query around `new_xyz`, neighbours' coordinates relative to their centre, optionally concatenated with their features."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        nbr = ball_query(self.radius, self.nsample, xyz, new_xyz)
        rel = grouping_operation(xyz.transpose(1, 2).contiguous(), nbr) - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return rel
        feats = grouping_operation(features, nbr)
        return torch.cat([rel, feats], dim=1) if self.use_xyz else feats


class GroupAll(nn.Module):                       # lib/pointnet2_utils.py:295-318
    """One group holding the whole cloud: (B,3,1,N) coordinates [+ (B,C,1,N) features]."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        whole = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return whole
        feats = features.unsqueeze(2)
        return torch.cat([whole, feats], dim=1) if self.use_xyz else feats
