"""The reference's self-supervised losses on the B200 kernels (losses/radar_loss.py:17-158): SoftChamferLoss, SpatialSmoothnessLoss,
RadialDisplacementLoss and their sum SelfSupervisedLoss -- same class names, constructor arguments, forward signatures and values,
differentiable with respect to the same inputs (pc1_warp / pred_flow).

The reference builds full (B,N,M) squared-distance matrices with torch.matmul and runs torch.topk over them
(utils/util.py:148-170).  Here the neighbour SEARCH is the library's brute-force k-NN kernel (cmf_knn_point: same expanded-form
ranking, O(N k) output, nothing materialised) and the kernel-density estimate is one reduction kernel (cmf_kde_density); both are
index / mask work through which no gradient flows in the reference either (topk indices, int32 masks).  The few values a gradient
does flow through -- the selected distances, the grouped flow vectors -- are recomputed on the selected pairs with ordinary torch
operations, so autograd gives the reference's gradients; the gather of the neighbours' flow vectors is the library's
grouping_operation, whose backward is cmf_group_points_grad.  There is no CPU path.
"""
import ctypes

import torch
from torch import nn
import torch.nn.functional as F

from . import pointnet2_utils as _pu
from ._lib import CmfError, check, dptr, lib, stream_ptr


def knn_point(k, xyz, new_xyz, with_dist=False):
    """radarflow_util.py:88-99 / utils/util.py:148-170 + topk(k, largest=False): indices (B,S,k) int32 of the k candidates of xyz (B,N,3)
    nearest to each query of new_xyz (B,S,3), ascending by (distance, index); optionally the clamped expanded-form squared distances."""
    if not xyz.is_cuda:
        raise CmfError("cmflow_b200.losses has no CPU path")
    xyz, new_xyz = xyz.detach().float().contiguous(), new_xyz.detach().float().contiguous()
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = torch.empty(B, S, k, dtype=torch.int32, device=xyz.device)
    dist = torch.empty(B, S, k, device=xyz.device) if with_dist else None
    with torch.cuda.device(xyz.device):
        check(lib().cmf_knn_point(B, N, S, k, dptr(xyz), dptr(new_xyz), dptr(idx), dptr(dist), stream_ptr()))
    return (idx, dist) if with_dist else idx


def kde_density(xyz1, xyz2, bandwidth):
    """compute_density_loss (utils/util.py:172-182): xyz1 (B,N,3), xyz2 (B,M,3) -> (B,N)."""
    xyz1, xyz2 = xyz1.detach().float().contiguous(), xyz2.detach().float().contiguous()
    B, N, _ = xyz1.shape
    out = torch.empty(B, N, device=xyz1.device)
    with torch.cuda.device(xyz1.device):
        check(lib().cmf_kde_density(B, N, xyz2.shape[1], dptr(xyz1), dptr(xyz2), ctypes.c_float(bandwidth), dptr(out), stream_ptr()))
    return out


def _selected_sqdist(a, b_sel):
    """The entries of square_distance(src, dst) (utils/util.py:166-169) at the selected pairs, with its arithmetic: -2 a.b + |a|^2 + |b|^2,
    clamped at zero.  a (B,N,3), b_sel (B,N,3)."""
    d = -2 * (a * b_sel).sum(-1) + (a ** 2).sum(-1) + (b_sel ** 2).sum(-1)
    return torch.clamp(d, min=0.0)


class SoftChamferLoss(nn.Module):
    """losses/radar_loss.py:17-60."""

    def __init__(self, zeta=0.005):
        super().__init__()
        self.zeta = zeta

    def forward(self, pc1, pc2, pc1_warp):
        p1 = pc1.permute(0, 2, 1).contiguous()                     # (B,N,3)
        p2 = pc2.permute(0, 2, 1).contiguous()
        pw = pc1_warp.permute(0, 2, 1).contiguous()
        mask1 = (kde_density(p1, p2, 1.0) > self.zeta).to(torch.int32)       # :39-43 inlier masks of the two clouds
        mask2 = (kde_density(p2, p1, 1.0) > self.zeta).to(torch.int32)
        # :45-50 nearest neighbour of every warped point in pc2 and of every pc2 point among the warped points; the min over a row / a column
        # of the distance matrix is the distance to that neighbour
        nn12 = knn_point(1, p2, pw).squeeze(2).long()               # (B,N): index into pc2
        nn21 = knn_point(1, pw, p2).squeeze(2).long()               # (B,M): index into pc1_warp
        d1 = _selected_sqdist(pw, torch.gather(p2, 1, nn12.unsqueeze(2).expand(-1, -1, 3)))
        d2 = _selected_sqdist(torch.gather(pw, 1, nn21.unsqueeze(2).expand(-1, -1, 3)), p2)
        d1 = F.relu(d1 - 0.01) * mask1                              # :53-56
        d2 = F.relu(d2 - 0.01) * mask2
        return torch.mean(d1) + torch.mean(d2)                      # :57


class SpatialSmoothnessLoss(nn.Module):
    """losses/radar_loss.py:62-98."""

    def __init__(self, alpha=0.5, num_nb=8):
        super().__init__()
        self.alpha = alpha
        self.num_nb = num_nb

    def forward(self, pc1, pred_flow):
        B, _, N = pc1.shape
        p1 = pc1.permute(0, 2, 1).contiguous()
        # :83-87 the num_nb + 1 nearest points sorted by distance, the first (the point itself) dropped
        kidx, dists = knn_point(self.num_nb + 1, p1, p1, with_dist=True)
        kidx, dists = kidx[:, :, 1:].contiguous(), dists[:, :, 1:]
        weights = torch.softmax(torch.exp(-dists / self.alpha).view(B, N * self.num_nb), dim=1).view(B, N, self.num_nb)    # :89-90
        # :92 index_points_group(pred_flow, kidx) = grouping_operation on the (B,3,N) flow (utils/util.py:52-63); backward: cmf_group_points_grad
        grouped = _pu.grouping_operation(pred_flow.contiguous(), kidx)          # (B,3,N,num_nb)
        diff = torch.norm(grouped - pred_flow.unsqueeze(3), dim=1)               # (B,N,num_nb)
        return torch.mean((N * weights * diff).sum(dim=2))                       # :93-94


class RadialDisplacementLoss(nn.Module):
    """losses/radar_loss.py:100-122 (the constructor argument is ignored there too: interval is 0.1)."""

    def __init__(self, intervel=0.1):
        super().__init__()
        self.interval = 0.1

    def forward(self, pc1, pred_f, vel1):
        pred_fr = torch.sum(pred_f * pc1, dim=1) / torch.norm(pc1, dim=1)
        return torch.mean(torch.abs(vel1 * self.interval - pred_fr))


class SelfSupervisedLoss(nn.Module):
    """losses/radar_loss.py:124-158."""

    def __init__(self, w_sc=1, w_ss=1, w_rd=1):
        super().__init__()
        self.w_sc, self.w_ss, self.w_rd = w_sc, w_ss, w_rd
        self.sc_loss, self.ss_loss, self.rd_loss = SoftChamferLoss(), SpatialSmoothnessLoss(), RadialDisplacementLoss()

    def forward(self, pc1, pc2, pred_f, vel1):
        scloss = self.sc_loss(pc1, pc2, pc1 + pred_f)
        ssloss = self.ss_loss(pc1, pred_f)
        rdloss = self.rd_loss(pc1, pred_f, vel1)
        total = self.w_sc * scloss + self.w_ss * ssloss + self.w_rd * rdloss
        return total, {"Loss": total.item(), "smoothnessLoss": ssloss.item(), "chamferLoss": scloss.item(), "veloLoss": rdloss.item()}
