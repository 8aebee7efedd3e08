"""Synthetic 4D-radar frame pairs and seeded CMFlow weights.

There is no View-of-Delft data and no network in the build environment, so bench.py, smoke()
and the parity fixtures all draw inputs from here.  Ranges follow the real radar clouds saved
under the reference's checkpoints/raflow_cvpr/results (x up to ~93 m, y ~ +-35 m)
and the feature layout of dataset/vod.py:62-63 ([RRV, RCS, RCS]).

`synthetic_state_dict` produces a state_dict with exactly the key layout of the reference's
checkpoints/cmflow_cvpr/models/model.best.t7 (374 keys; cmflow_t adds the four GRU tensors), with
He-style conv weights and non-trivial BatchNorm running statistics, from a name-independent seeded
stream -- so the fixture generator (which feeds it to the *reference* model) and the tests (which feed
it to this package) get bit-identical weights without committing a 17 MB checkpoint.
"""
import math

import torch

SA_RADIUS = (2.0, 4.0, 8.0, 16.0)     # models/cmflow.py:21,35
SA_NSAMPLE = (4, 8, 16, 32)           # models/cmflow.py:22,36


def make_pairs(B, N, seed=1234, dense=False, device="cpu"):
    """Returns pc1, pc2, ft1, ft2, each (B,3,N) float32 contiguous, plus gt transform (B,4,4).

    pc2 = R_gt pc1 + t_gt + noise, then an independent permutation of pc2's points so index order
    carries no correspondence.  dense=True uses the N=4096 "LiDAR stress" ranges (BASELINE.json configs[4]).
    """
    g = torch.Generator().manual_seed(seed)
    xr, yr = (80.0, 40.0) if dense else (50.0, 20.0)
    x = torch.rand(B, N, generator=g) * xr
    y = (torch.rand(B, N, generator=g) * 2 - 1) * yr
    z = (torch.rand(B, N, generator=g) * 2 - 1) * 2.0
    pc1 = torch.stack([x, y, z], 1)                                   # (B,3,N)
    yaw = (torch.rand(B, generator=g) * 2 - 1) * math.radians(1.0)
    t = torch.stack([torch.rand(B, generator=g) * 1.5, (torch.rand(B, generator=g) * 2 - 1) * 0.1,
                     torch.zeros(B)], 1)
    c, s = torch.cos(yaw), torch.sin(yaw)
    R = torch.zeros(B, 3, 3)
    R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1], R[:, 2, 2] = c, -s, s, c, 1.0
    pc2 = torch.bmm(R, pc1) + t[:, :, None] + torch.randn(B, 3, N, generator=g) * 0.05
    perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])  # (B,N)
    pc2 = torch.gather(pc2, 2, perm[:, None, :].expand(B, 3, N))

    def feats():
        rrv = torch.randn(B, 1, N, generator=g) * 2.0
        rcs = (torch.rand(B, 1, N, generator=g) * 2 - 1) * 20.0
        return torch.cat([rrv, rcs, rcs], 1)

    ft1, ft2 = feats(), feats()
    T = torch.eye(4).repeat(B, 1, 1)
    T[:, :3, :3], T[:, :3, 3] = R, t
    out = [v.contiguous().float().to(device) for v in (pc1, pc2, ft1, ft2, T)]
    return tuple(out)


def make_padded_pairs(B, N, n_unique, seed):
    """Pairs whose clouds hold n_unique distinct points padded to N by duplicate sampling, exactly as the training loader does
    (dataset/vod.py:102-110: arange(npts) followed by np.random.choice(npts, N - npts, replace=True); positions AND features are
    indexed with the same sample).  Returns the pairs and the (B,N) maps point -> unique id of each cloud."""
    pc1, pc2, ft1, ft2, T = make_pairs(B, n_unique, seed=seed)
    g = torch.Generator().manual_seed(seed + 7)
    out, maps = [], []
    for pc, ft in ((pc1, ft1), (pc2, ft2)):
        idx = torch.stack([torch.cat([torch.arange(n_unique), torch.randint(0, n_unique, (N - n_unique,), generator=g)]) for _ in range(B)])
        gi = idx[:, None, :].expand(B, 3, N)
        out.append((torch.gather(pc, 2, gi).contiguous(), torch.gather(ft, 2, gi).contiguous()))
        maps.append(idx)
    return (out[0][0], out[1][0], out[0][1], out[1][1], T), maps, (pc1, pc2)


def _conv(g, cout, cin, bias, gain=1.0):
    # std = gain / sqrt(fan_in): keeps activations O(1) through the ~20 stacked layers so that the
    # motion scores are spread around the 0.5 threshold and the Kabsch system is well conditioned.
    d = {"weight": torch.randn(cout, cin, 1, 1, generator=g) * (gain / math.sqrt(cin))}
    if bias:
        d["bias"] = torch.randn(cout, generator=g) * 0.1
    return d


def _bn(g, c):
    return {"weight": torch.rand(c, generator=g) + 0.5, "bias": torch.randn(c, generator=g) * 0.1,
            "running_mean": torch.randn(c, generator=g) * 0.2, "running_var": torch.rand(c, generator=g) + 0.5,
            "num_batches_tracked": torch.tensor(1000, dtype=torch.int64)}


def synthetic_state_dict(seed=0, temporal=False):
    """Key layout of models/cmflow.py:12-48 (+ cmflow_t.py:46 GRU when temporal)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def put(prefix, d):
        for k, v in d.items():
            sd[prefix + "." + k] = v

    def set_conv(prefix, cin, mlp, mlp2):          # PointLocalFeature, radarflow_util.py:121-142
        last = cin + 3
        for i, co in enumerate(mlp):
            put(f"{prefix}.mlp_convs.{i}", _conv(g, co, last, False, gain=(0.25 if (i == 0 and cin == 3) else 1.4)))
            put(f"{prefix}.mlp_bns.{i}", _bn(g, co))
            last = co
        for i, co in enumerate(mlp2):
            put(f"{prefix}.mlp2_convs.{i}", _conv(g, co, last, False, gain=1.4))
            put(f"{prefix}.mlp2_bns.{i}", _bn(g, co))
            last = co

    def weightnet(prefix):                           # WeightNet, radarflow_util.py:288-305 (BNs exist, unused)
        dims = [(8, 3), (8, 8), (512, 8)]
        for i, (co, ci) in enumerate(dims):
            put(f"{prefix}.mlp_convs.{i}", _conv(g, co, ci, True, gain=(0.2 if i == 0 else 0.7)))
            put(f"{prefix}.mlp_bns.{i}", _bn(g, co))

    def head(prefix, cout):                          # FlowHead / MotionHead, radarflow_util.py:240-285
        last = 512
        for i, co in enumerate((256, 128, 64)):
            put(f"{prefix}.sf_mlp.{i}.0", _conv(g, co, last, False, gain=1.4))
            put(f"{prefix}.sf_mlp.{i}.1", _bn(g, co))
            last = co
        c2 = _conv(g, cout, 64, False, gain=(1.0 if cout == 1 else 0.5))
        if cout == 1:
            # the last trunk layer is post-ReLU (all channels >= 0): a zero-mean, low-variance read-out keeps the
            # motion logits spread around 0 instead of saturating with one sign for the whole cloud
            c2["weight"] = (c2["weight"] - c2["weight"].mean()) * 2.0
            # ... and damp the per-cloud constant (global-feature columns 256:512 of the first trunk conv)
            sd[f"{prefix}.sf_mlp.0.0.weight"][:, 256:] *= 0.05
        put(f"{prefix}.conv2", c2)

    for l in range(4):
        set_conv(f"mse_layer.ms_ls.{l}", 3, (32, 32, 64), (64, 64, 64))
    last = 1027
    for i in range(3):
        put(f"fc_layer.mlp_convs.{i}", _conv(g, 512, last, True, gain=1.3))
        last = 512
    weightnet("fc_layer.weightnet1")
    weightnet("fc_layer.weightnet2")
    for l in range(4):
        set_conv(f"mse_layer2.ms_ls.{l}", 1027, (512, 256, 64), (64, 64, 64))
    if temporal:
        k = 1.0 / math.sqrt(256)
        for name, shape in (("weight_ih_l0", (768, 256)), ("weight_hh_l0", (768, 256)),
                            ("bias_ih_l0", (768,)), ("bias_hh_l0", (768,))):
            sd["gru." + name] = (torch.rand(*shape, generator=g) * 2 - 1) * k
    head("fp", 3)
    head("mp", 1)
    return sd


def raflow_state_dict(seed=0):
    """Key layout of models/raflow.py:11-35: the CMFlow backbone with the decoder under fd_layer (FlowDecoder,
    radarflow_util.py:321-337: fd_layer.mse = MultiScaleEncoder(1027 -> 512/256/64), fd_layer.fp = FlowPredictor) and no motion head."""
    out = {}
    for k, v in synthetic_state_dict(seed).items():
        if k.startswith("mp."):
            continue
        if k.startswith("mse_layer2."):
            k = "fd_layer.mse." + k[len("mse_layer2."):]
        elif k.startswith("fp."):
            k = "fd_layer.fp." + k[len("fp."):]
        out[k] = v
    return out
