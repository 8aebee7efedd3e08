"""CPU restatement of CMFlow / CMFlow-T inference (TEST INFRASTRUCTURE -- see oracle/__init__.py).

A functional, direct-form (un-hoisted, un-fused) restatement of the reference forward pass that
takes the reference's own state_dict.  Every stage cites the reference lines it follows (paths
relative to the upstream tree).  `dtype=torch.float64` evaluates all real arithmetic in double --
while neighbour indices are still selected with the reference's float32 arithmetic -- which gives a
"truth" against which both the reference's fp32 result and the CUDA result can be placed.

Parity pin: tests/golden/*.pt are produced by tests/golden/make_golden.py, which runs the UNMODIFIED
reference Python (models/cmflow.py, utils/model_utils/radarflow_util.py, lib/pointnet2_utils.py)
on CPU; tests/test_oracle_golden.py checks this file against them.  Third-party arithmetic the
reference delegates to torch (conv/BN/matmul/topk/svd/GRU) is whatever the installed torch computes
in fp32 on CPU; no reference test pins those call sites.
"""
import torch

from . import pointops as P

SA_RADIUS = (2.0, 4.0, 8.0, 16.0)      # models/cmflow.py:21,35
SA_NSAMPLE = (4, 8, 16, 32)            # models/cmflow.py:22,36
BN_EPS = 1e-5                          # nn.BatchNorm2d default, radarflow_util.py:133


def _conv(sd, name, x, dtype):
    """1x1 Conv2d over a (B,C,...) tensor: radarflow_util.py:132,138,174,246."""
    w = sd[name + ".weight"].to(dtype)[:, :, 0, 0]
    y = torch.einsum("oc,bc...->bo...", w, x)
    b = sd.get(name + ".bias")
    if b is not None:
        y = y + b.to(dtype).view(1, -1, *([1] * (x.dim() - 2)))
    return y


def _bn(sd, name, x, dtype):
    """BatchNorm2d in eval mode (running statistics)."""
    sh = (1, -1) + (1,) * (x.dim() - 2)
    m, v = sd[name + ".running_mean"].to(dtype), sd[name + ".running_var"].to(dtype)
    g, b = sd[name + ".weight"].to(dtype), sd[name + ".bias"].to(dtype)
    return (x - m.view(sh)) / torch.sqrt(v.view(sh) + BN_EPS) * g.view(sh) + b.view(sh)


def _group(feat, idx):
    """grouping_operation (pointnet2_utils.py:187-205): feat (B,C,N), idx (B,P,S) -> (B,C,P,S)."""
    B, C, N = feat.shape
    _, Pn, S = idx.shape
    ix = idx.long().view(B, 1, Pn * S).expand(B, C, Pn * S)
    return torch.gather(feat, 2, ix).view(B, C, Pn, S)


def point_local_feature(sd, prefix, radius, nsample, xyz, feats, dtype, idx=None):
    """PointLocalFeature.forward (radarflow_util.py:144-162) incl. QueryAndGroup (pointnet2_utils.py:269-292).
    xyz (B,3,N) f32, feats (B,C,N) -> (B,64,N)."""
    xyz_t = xyz.permute(0, 2, 1).contiguous()
    if idx is None:
        idx = P.ball_query(radius, nsample, xyz_t, xyz_t)                 # :277, float32 always
    g_xyz = _group(xyz.to(dtype), idx) - xyz.to(dtype).unsqueeze(-1)       # :279-280
    x = torch.cat([g_xyz, _group(feats, idx)], 1)                          # :283-285
    for i in range(3):                                                     # :151-153
        x = torch.relu(_bn(sd, f"{prefix}.mlp_bns.{i}", _conv(sd, f"{prefix}.mlp_convs.{i}", x, dtype), dtype))
    x = x.max(-1)[0].unsqueeze(2)                                          # :155
    for i in range(3):                                                     # :157-159
        x = torch.relu(_bn(sd, f"{prefix}.mlp2_bns.{i}", _conv(sd, f"{prefix}.mlp2_convs.{i}", x, dtype), dtype))
    return x.squeeze(2)


def multi_scale_encoder(sd, prefix, xyz, feats, dtype, idxs=None):
    """MultiScaleEncoder.forward (radarflow_util.py:111-118): concat of the 4 scales on channels."""
    outs = [point_local_feature(sd, f"{prefix}.ms_ls.{l}", SA_RADIUS[l], SA_NSAMPLE[l], xyz, feats, dtype,
                                None if idxs is None else idxs[l]) for l in range(4)]
    return torch.cat(outs, 1)


def _weightnet(sd, prefix, d, dtype):
    """WeightNet.forward with bn=False (radarflow_util.py:307-318): ReLU after every conv."""
    for i in range(3):
        d = torch.relu(_conv(sd, f"{prefix}.mlp_convs.{i}", d, dtype))
    return d


def feature_correlator(sd, xyz1, xyz2, pts1, pts2, dtype, knn12=None, knn11=None):
    """FeatureCorrelator.forward (radarflow_util.py:185-237). xyz (B,3,N) f32, pts (B,D,N) -> (B,512,N)."""
    B, _, N1 = xyz1.shape
    x1t = xyz1.permute(0, 2, 1).contiguous()
    x2t = xyz2.permute(0, 2, 1).contiguous()
    if knn12 is None:
        knn12 = P.knn_point(8, x2t, x1t)[0]                                # :207 (float32 always)
    nb_xyz = _group(xyz2.to(dtype), knn12)                                 # (B,3,N1,8)   :208
    direction = nb_xyz - xyz1.to(dtype).unsqueeze(-1)                      # :209
    g2 = _group(pts2, knn12)                                               # :211
    g1 = pts1.unsqueeze(-1).expand(-1, -1, -1, 8)                          # :212
    x = torch.cat([g1, g2, direction], 1)                                  # :213-214 (channel order D1, D2, 3)
    for i in range(3):                                                     # :215-220 LeakyReLU(0.1), bias, no bn
        x = torch.nn.functional.leaky_relu(_conv(sd, f"fc_layer.mlp_convs.{i}", x, dtype), 0.1)
    w = _weightnet(sd, "fc_layer.weightnet1", direction, dtype)            # :223
    p2p = (w * x).sum(-1)                                                  # :225  (B,512,N1)
    if knn11 is None:
        knn11 = P.knn_point(8, x1t, x1t)[0]                                # :228
    nb_xyz = _group(xyz1.to(dtype), knn11)
    direction = nb_xyz - xyz1.to(dtype).unsqueeze(-1)                      # :230
    w = _weightnet(sd, "fc_layer.weightnet2", direction, dtype)            # :233
    return (w * _group(p2p, knn11)).sum(-1)                                # :234-235


def _head(sd, prefix, x, dtype):
    """FlowHead / MotionHead trunk (radarflow_util.py:240-285)."""
    for i in range(3):
        x = torch.relu(_bn(sd, f"{prefix}.sf_mlp.{i}.1", _conv(sd, f"{prefix}.sf_mlp.{i}.0", x, dtype), dtype))
    return _conv(sd, f"{prefix}.conv2", x, dtype)


def weighted_kabsch(A, Bp, W):
    """CMFlow.WeightedKabsch (models/cmflow.py:128-169). A,Bp (B,3,N); W (B,N) -> (B,4,4).
    Includes the reference's reflection handling: ROW 2 of V is negated (cmflow.py:162)."""
    Wc = W.unsqueeze(2)
    cA = (A.transpose(2, 1) * Wc).sum(1).unsqueeze(2)                      # :138
    cB = (Bp.transpose(2, 1) * Wc).sum(1).unsqueeze(2)                     # :139
    Am, Bm = A - cA, Bp - cB                                               # :148-149
    H = Am @ (Bm.transpose(2, 1) * Wc)                                     # :151
    U, _, Vh = torch.linalg.svd(H)                                         # :154 (torch.svd returns V = Vh^T)
    V = Vh.transpose(2, 1)
    Z = V @ U.transpose(2, 1)                                              # :155
    d = (torch.linalg.det(Z) < 0).to(A.dtype) * 2 - 1                      # :157-160  (+1 when reflected)
    Vc = V.clone()
    Vc[:, 2, :] = Vc[:, 2, :] * (-d.view(-1, 1))                           # :162
    R = Vc @ U.transpose(2, 1)                                             # :163
    t = -R @ cA + cB                                                       # :165
    T = torch.zeros(A.shape[0], 4, 4, dtype=A.dtype, device=A.device)
    T[:, :3, :3], T[:, :3, 3:], T[:, 3, 3] = R, t, 1.0                     # :167
    return T, H


def rigid_to_flow(pc, T):
    """CMFlow.rigid_to_flow (models/cmflow.py:51-55)."""
    h = torch.cat([pc, torch.ones(pc.shape[0], 1, pc.shape[2], dtype=pc.dtype, device=pc.device)], 1)
    return (T @ h)[:, :3] - pc


def gru_step(sd, x, h, dtype):
    """One nn.GRU(256,256) step (models/cmflow_t.py:46,101): gates ordered r,z,n."""
    Wi, Wh = sd["gru.weight_ih_l0"].to(dtype), sd["gru.weight_hh_l0"].to(dtype)
    bi, bh = sd["gru.bias_ih_l0"].to(dtype), sd["gru.bias_hh_l0"].to(dtype)
    gi, gh = x @ Wi.t() + bi, h @ Wh.t() + bh
    ir, iz, in_ = gi.chunk(3, 1)
    hr, hz, hn = gh.chunk(3, 1)
    r, z = torch.sigmoid(ir + hr), torch.sigmoid(iz + hz)
    n = torch.tanh(in_ + r * hn)
    return (1 - z) * n + z * h


def cmflow_forward(sd, pc1, pc2, ft1, ft2, stat_thres=0.5, dtype=torch.float32, temporal=False, gfeat_prev=None,
                   return_intermediates=False, label_m=None):
    """CMFlow.forward(pc1,pc2,feature1,feature2,label_m=None,mode='test') (models/cmflow.py:171-197) and, with
    temporal=True, CMFlow_T.forward(..., gfeat) (models/cmflow_t.py:185-211).

    Inputs (B,3,N) float32.  Returns dict with sf_agg (B,3,N), stat_cls (B,1,N), pre_trans (B,4,4), mask (B,N) bool
    [, gfeat (B,256)] in `dtype`, plus intermediates when asked."""
    pc1, pc2 = pc1.float().contiguous(), pc2.float().contiguous()
    f1in, f2in = ft1.to(dtype), ft2.to(dtype)
    B, _, N = pc1.shape
    x1t = pc1.permute(0, 2, 1).contiguous()
    x2t = pc2.permute(0, 2, 1).contiguous()
    bq1 = [P.ball_query(SA_RADIUS[l], SA_NSAMPLE[l], x1t, x1t) for l in range(4)]
    bq2 = [P.ball_query(SA_RADIUS[l], SA_NSAMPLE[l], x2t, x2t) for l in range(4)]
    # Backbone, cmflow.py:59-93
    f1 = multi_scale_encoder(sd, "mse_layer", pc1, f1in, dtype, bq1)          # :72
    f2 = multi_scale_encoder(sd, "mse_layer", pc2, f2in, dtype, bq2)          # :73
    g1 = f1.max(-1)[0].unsqueeze(2).expand(-1, -1, N)                         # :76
    g2 = f2.max(-1)[0].unsqueeze(2).expand(-1, -1, pc2.shape[2])              # :77
    pf1, pf2 = torch.cat([f1, g1], 1), torch.cat([f2, g2], 1)                 # :80-81
    knn12 = P.knn_point(8, x2t, x1t)[0]
    knn11 = P.knn_point(8, x1t, x1t)[0]
    cor = feature_correlator(sd, pc1, pc2, pf1, pf2, dtype, knn12, knn11)     # :84
    emb = torch.cat([f1in, pf1, cor], 1)                                      # :87
    prop = multi_scale_encoder(sd, "mse_layer2", pc1, emb, dtype, bq1)        # :88
    gfeat = prop.max(-1)[0]                                                   # :89
    gnew = None
    if temporal:                                                              # cmflow_t.py:94-105
        h0 = torch.zeros_like(gfeat) if gfeat_prev is None else gfeat_prev.to(dtype)
        gnew = gru_step(sd, gfeat, h0, dtype)
        gexp = gnew.unsqueeze(2).expand(-1, -1, N)
    else:
        gexp = gfeat.unsqueeze(2).expand(-1, -1, N)
    final = torch.cat([prop, gexp], 1)                                        # :91
    flow = _head(sd, "fp", final, dtype)                                      # :177
    stat_cls = torch.sigmoid(_head(sd, "mp", final, dtype))                   # :178
    # mode='train' with pseudo labels: scores = label_m.unsqueeze(1) (:181-182); otherwise the predicted scores (:184-185)
    scores = stat_cls if label_m is None else label_m.to(dtype).unsqueeze(1)
    mask = (scores > stat_thres).squeeze(1)                                   # :188
    score = scores.squeeze(1)
    if not temporal:
        score = score + 1e-4                                                  # cmflow.py:105 (absent in cmflow_t.py:119)
    weight = score / score.sum(1, keepdim=True)                               # :106
    pcd = pc1.to(dtype)
    T, H = weighted_kabsch(pcd, pcd + flow, weight)                           # :108
    sf_rg = rigid_to_flow(pcd, T)                                             # :116
    sf_agg = torch.where(mask.unsqueeze(1), sf_rg, flow)                      # :119-123 (per-sample masked copies)
    out = {"sf_agg": sf_agg, "stat_cls": stat_cls, "pre_trans": T, "mask": mask}
    if temporal:
        out["gfeat"] = gnew
    if return_intermediates:
        out.update({"bq1": bq1, "bq2": bq2, "knn12": knn12, "knn11": knn11, "f1": f1, "f2": f2, "cor": cor,
                    "prop": prop, "flow": flow, "H": H, "weight": weight})
    return out


# ---- RaFlow (models/raflow.py) -------------------------------------------------------------------------------------------------
def raflow_rigid_transform(A, Bp, M):
    """RaFlow.rigid_transform_torch (models/raflow.py:119-156). A,Bp (B,3,N); M (B,N) 0/1 mask -> (B,4,4).
    Quirks kept: the centroids are torch.mean over ALL N points of the masked coordinates (raflow.py:129-130: divided by N, not
    by the number of masked points); every point is centred (:137-138) but only masked columns enter H (:140); reflection
    handling negates ROW 2 of V (:151)."""
    W = M.to(torch.bool).unsqueeze(2).to(A.dtype)
    cA = (A.transpose(2, 1) * W).mean(1).unsqueeze(2)                       # :129,133
    cB = (Bp.transpose(2, 1) * W).mean(1).unsqueeze(2)                      # :130,134
    Am, Bm = A - cA, Bp - cB                                               # :137-138
    H = Am @ (Bm.transpose(2, 1) * W)                                      # :140
    U, _, Vh = torch.linalg.svd(H)                                         # :143
    V = Vh.transpose(2, 1)
    Z = V @ U.transpose(2, 1)                                              # :144
    d = (torch.linalg.det(Z) < 0).to(A.dtype) * 2 - 1                      # :146-149
    Vc = V.clone()
    Vc[:, 2, :] = Vc[:, 2, :] * (-d.view(-1, 1))                           # :151
    R = Vc @ U.transpose(2, 1)                                             # :152
    t = -R @ cA + cB                                                       # :154
    T = torch.zeros(A.shape[0], 4, 4, dtype=A.dtype, device=A.device)
    T[:, :3, :3], T[:, :3, 3:], T[:, 3, 3] = R, t, 1.0                     # :156
    return T


def raflow_forward(sd, pc1, pc2, ft1, ft2, interval, rigid_thres=0.15, rigid_pcs=0.25, dtype=torch.float32,
                   return_intermediates=False):
    """RaFlow.forward(pc1, pc2, feature1, feature2, interval) (models/raflow.py:157-164) -> dict with output (B,3,N),
    sf_agg (B,3,N), pre_trans (B,4,4), mask_s (B,N) bool.  interval (B,)."""
    pc1, pc2 = pc1.float().contiguous(), pc2.float().contiguous()
    f1in, f2in = ft1.to(dtype), ft2.to(dtype)
    B, _, N = pc1.shape
    x1t = pc1.permute(0, 2, 1).contiguous()
    x2t = pc2.permute(0, 2, 1).contiguous()
    bq1 = [P.ball_query(SA_RADIUS[l], SA_NSAMPLE[l], x1t, x1t) for l in range(4)]
    bq2 = [P.ball_query(SA_RADIUS[l], SA_NSAMPLE[l], x2t, x2t) for l in range(4)]
    # ROFE_module, raflow.py:47-77
    f1 = multi_scale_encoder(sd, "mse_layer", pc1, f1in, dtype, bq1)          # :59
    f2 = multi_scale_encoder(sd, "mse_layer", pc2, f2in, dtype, bq2)          # :60
    g1 = f1.max(-1)[0].unsqueeze(2).expand(-1, -1, N)                         # :63
    g2 = f2.max(-1)[0].unsqueeze(2).expand(-1, -1, pc2.shape[2])              # :64
    pf1, pf2 = torch.cat([f1, g1], 1), torch.cat([f2, g2], 1)                 # :67-68
    knn12 = P.knn_point(8, x2t, x1t)[0]
    knn11 = P.knn_point(8, x1t, x1t)[0]
    cor = feature_correlator(sd, pc1, pc2, pf1, pf2, dtype, knn12, knn11)     # :71
    # FlowDecoder.forward, radarflow_util.py:339-350
    emb = torch.cat([f1in, pf1, cor], 1)                                      # :341
    prop = multi_scale_encoder(sd, "fd_layer.mse", pc1, emb, dtype, bq1)      # :343
    gfeat = prop.max(-1)[0].unsqueeze(2).expand(-1, -1, N)                    # :344
    output = _head(sd, "fd_layer.fp", torch.cat([prop, gfeat], 1), dtype)     # :345-348 (FlowPredictor :388-409)
    # SFR_module, raflow.py:79-117
    pcd = pc1.to(dtype)
    warp = pcd + output                                                       # :85
    ones = torch.ones(B, N, dtype=dtype, device=pc1.device)
    trans = raflow_rigid_transform(pcd, warp, ones)                           # :89-90
    sf_rg = rigid_to_flow(pcd, trans)                                         # :92
    vel = f1in[:, 0]                                                          # :95
    sf_proj = (sf_rg * pcd).sum(1) / torch.norm(pcd, dim=1)                   # :96
    residual = vel * interval.to(dtype).unsqueeze(1) - sf_proj                # :97
    mask_s = (residual / vel).abs() < rigid_thres                             # :98
    pre_trans = torch.zeros_like(trans)
    sf_agg = torch.zeros_like(output)
    for b in range(B):                                                        # :103-114
        if (mask_s[b].sum() / N) > rigid_pcs:
            pre_trans[b] = raflow_rigid_transform(pcd[b:b + 1], warp[b:b + 1], mask_s[b:b + 1])[0]
            sf_agg[b] = rigid_to_flow(pcd[b:b + 1], pre_trans[b:b + 1])[0]
            keep = torch.logical_not(mask_s[b])
            sf_agg[b][:, keep] = output[b][:, keep]
        else:
            pre_trans[b] = trans[b]
            sf_agg[b] = output[b]
    out = {"output": output, "sf_agg": sf_agg, "pre_trans": pre_trans, "mask_s": mask_s}
    if return_intermediates:
        out.update({"f1": f1, "f2": f2, "cor": cor, "prop": prop, "trans0": trans, "residual": residual, "vel": vel})
    return out
