"""CPU emulation of csrc/engine.cu's launch sequence on the PACKED blob (test-only).

Purpose: validate, without a GPU, (a) cmflow_b200/weights.py (BN folding, column-block splits, segment
order) and (b) the hoisting algebra of the engine, by replaying forward_chunk() stage by stage with torch
CPU ops on exactly the segments the CUDA engine reads, and comparing against the oracle.  Neighbour
indices come from the C oracle (the CUDA kernels are compared with it separately, bit-exactly).
"""
import numpy as np
import torch

from oracle import pointops as P

KS = (4, 8, 16, 32)
KOFF = (0, 4, 12, 28)
E_LD = 800


def read_segments(blob):
    hdr = blob[:512].view(np.int32)
    nseg = int(hdr[1])
    segs = []
    for i in range(nseg):
        off, r, c = (int(v) for v in hdr[3 + 3 * i: 6 + 3 * i])
        segs.append(torch.from_numpy(blob[off: off + r * c].reshape(r, c).copy()))
    return segs


def gemm(W, X, bias=None, pbias=None, cpp=1, act=None, dtype=torch.float32):
    Y = X.to(dtype) @ W.to(dtype).t()
    if bias is not None:
        Y = Y + bias.to(dtype).view(1, -1)
    if pbias is not None:
        Y = Y + pbias.to(dtype).repeat_interleave(cpp, 0)
    if act == "relu":
        Y = torch.relu(Y)
    elif act == "leaky":
        Y = torch.nn.functional.leaky_relu(Y, 0.1)
    return Y


def emulate(blob, pc1, pc2, ft1, ft2, temporal=False, gprev=None, stat_thres=0.5, dtype=torch.float32):
    S = read_segments(blob)
    B, _, N = pc1.shape
    BN = B * N
    x1t = pc1.permute(0, 2, 1).contiguous(); x2t = pc2.permute(0, 2, 1).contiguous()
    bq = {}
    for name, xt in (("1", x1t), ("2", x2t)):
        bq[name] = torch.cat([P.ball_query(r, k, xt, xt) for r, k in zip((2.0, 4.0, 8.0, 16.0), KS)], -1).long()   # (B,N,60)
    knn12 = P.knn_point(8, x2t, x1t)[0].long(); knn11 = P.knn_point(8, x1t, x1t)[0].long()
    bidx = torch.arange(B).view(B, 1, 1)

    def rows(pl):                       # (B,3,N) planar -> (B,N,3)
        return pl.permute(0, 2, 1).to(dtype)

    def mse_layer(pc, ft, idx60):
        xyz, f = rows(pc), rows(ft)
        outs = []
        for s in range(4):
            sb = s * 12
            j = idx60[:, :, KOFF[s]:KOFF[s] + KS[s]]                                   # (B,N,K)
            rel = xyz[bidx, j] - xyz[:, :, None, :]
            x0 = torch.cat([rel, f[bidx, j], torch.zeros(B, N, KS[s], 2, dtype=dtype)], -1).reshape(-1, 8)
            t = gemm(S[sb + 0], x0, S[sb + 1][0], act="relu", dtype=dtype)
            t = gemm(S[sb + 2], t, S[sb + 3][0], act="relu", dtype=dtype)
            t = gemm(S[sb + 4], t, S[sb + 5][0], act="relu", dtype=dtype)
            m = t.view(BN, KS[s], 64).max(1)[0]
            for l in range(3):
                m = gemm(S[sb + 6 + 2 * l], m, S[sb + 7 + 2 * l][0], act="relu", dtype=dtype)
            outs.append(m)
        F = torch.cat(outs, 1)                                                          # (BN,256)
        return F, F.view(B, N, 256).max(1)[0]

    F1, G1 = mse_layer(pc1, ft1, bq["1"])
    F2, G2 = mse_layer(pc2, ft2, bq["2"])
    FC = 48
    PB1 = gemm(S[FC + 1], G1, S[FC + 5][0], dtype=dtype); PB2 = gemm(S[FC + 3], G2, dtype=dtype)
    U1 = gemm(S[FC + 0], F1, pbias=PB1, cpp=N, dtype=dtype).view(B, N, 512)
    U2 = gemm(S[FC + 2], F2, pbias=PB2, cpp=N, dtype=dtype).view(B, N, 512)
    xyz1, xyz2 = rows(pc1), rows(pc2)
    d12 = xyz2[bidx, knn12] - xyz1[:, :, None, :]                                        # (B,N,8,3)
    Wd = S[FC + 4].to(dtype)[:, :3]
    H1 = torch.nn.functional.leaky_relu(U1[:, :, None, :] + U2[bidx, knn12] + d12 @ Wd.t(), 0.1).reshape(-1, 512)
    H2 = gemm(S[FC + 6], H1, S[FC + 7][0], act="leaky", dtype=dtype)
    H3 = gemm(S[FC + 8], H2, S[FC + 9][0], act="leaky", dtype=dtype).view(B, N, 8, 512)

    def weightnet(base, d):
        h = torch.relu(d @ S[base].to(dtype)[:, :3].t() + S[base + 1].to(dtype))
        h = torch.relu(h @ S[base + 2].to(dtype).t() + S[base + 3].to(dtype))
        return torch.relu(h @ S[base + 4].to(dtype).t() + S[base + 5].to(dtype))

    cost1 = (weightnet(58, d12) * H3).sum(2)                                             # (B,N,512)
    d11 = xyz1[bidx, knn11] - xyz1[:, :, None, :]
    cor = (weightnet(64, d11) * cost1[bidx, knn11]).sum(2)
    E = torch.zeros(B, N, E_LD, dtype=dtype)
    E[:, :, 0:256] = F1.view(B, N, 256); E[:, :, 256:768] = cor; E[:, :, 768:771] = rows(ft1)
    PBM = gemm(S[71], G1, S[72][0], dtype=dtype)
    Pm = gemm(S[70], E.view(BN, E_LD), pbias=PBM, cpp=N, dtype=dtype).view(B, N, 2048)
    outs = []
    for s in range(4):
        sb = 74 + s * 10
        j = bq["1"][:, :, KOFF[s]:KOFF[s] + KS[s]]
        rel = xyz1[bidx, j] - xyz1[:, :, None, :]
        Wx = S[73].to(dtype)[s * 512:(s + 1) * 512, :3]
        y1 = torch.relu(Pm[bidx, j][..., s * 512:(s + 1) * 512] + rel @ Wx.t()).reshape(-1, 512)
        y2 = gemm(S[sb], y1, S[sb + 1][0], act="relu", dtype=dtype)
        y3 = gemm(S[sb + 2], y2, S[sb + 3][0], act="relu", dtype=dtype)
        m = y3.view(BN, KS[s], 64).max(1)[0]
        for l in range(3):
            m = gemm(S[sb + 4 + 2 * l], m, S[sb + 5 + 2 * l][0], act="relu", dtype=dtype)
        outs.append(m)
    PROP = torch.cat(outs, 1)
    GP = PROP.view(B, N, 256).max(1)[0]
    gvec, gnew = GP, None
    if temporal:
        gi = gemm(S[126], GP, S[128][0], dtype=dtype)
        hp = torch.zeros(B, 256, dtype=dtype) if gprev is None else gprev.to(dtype)
        gh = gemm(S[127], hp, S[129][0], dtype=dtype)
        r = torch.sigmoid(gi[:, :256] + gh[:, :256]); z = torch.sigmoid(gi[:, 256:512] + gh[:, 256:512])
        nn_ = torch.tanh(gi[:, 512:] + r * gh[:, 512:])
        gnew = (1 - z) * nn_ + z * hp
        gvec = gnew
    PBH = gemm(S[115], gvec, S[116][0], dtype=dtype)
    HD1 = gemm(S[114], PROP, pbias=PBH, cpp=N, act="relu", dtype=dtype)
    f2 = gemm(S[117], HD1[:, :256], S[118][0], act="relu", dtype=dtype); m2 = gemm(S[119], HD1[:, 256:], S[120][0], act="relu", dtype=dtype)
    f3 = gemm(S[121], f2, S[122][0], act="relu", dtype=dtype); m3 = gemm(S[123], m2, S[124][0], act="relu", dtype=dtype)
    W4 = S[125].to(dtype)
    flow = (f3 @ W4[:3].t()).view(B, N, 3).permute(0, 2, 1)
    cls = torch.sigmoid(m3 @ W4[3:4].t()).view(B, 1, N)
    return {"f1": F1.view(B, N, 256), "f2": F2.view(B, N, 256), "cor": cor, "prop": PROP.view(B, N, 256), "flow": flow,
            "stat_cls": cls, "gfeat": gnew, "knn12": knn12, "knn11": knn11, "bq1": bq["1"], "bq2": bq["2"]}
