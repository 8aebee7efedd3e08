"""Pack a reference-layout state_dict into the engine's weight blob (the contract with csrc/engine.cu).

All folding is done in float64 on the host, then rounded once to float32:
  * eval-mode BatchNorm2d after a bias-free 1x1 conv:  y = s*(W x) + t,  s = gamma/sqrt(var+eps), t = beta - s*mean
    -> W' = s[:,None]*W, bias t                         (radarflow_util.py:151-153,157-159,246-252)
  * conv over a channel concat is split by column block (see the algebra note at the top of engine.cu):
      fc_layer.mlp_convs.0  (512 x 1027): [f1 256 | g1 256 | f2 256 | g2 256 | dir 3]   (radarflow_util.py:213)
      mse_layer2 conv0      (512 x 1030): [rel-xyz 3 | ft 3 | f1 256 | g1 256 | cor 512] (pointnet2_utils.py:285, cmflow.py:87)
      head conv0            (256 x 512):  [prop 256 | gfeat 256]                         (cmflow.py:91)

Blob = 512-float header (int32 view: magic, nseg, temporal, then (offset, rows, cols) per segment) followed by
the segments, each padded to a multiple of 4 floats.  Segment order == the enum in csrc/engine.cu.
"""
import numpy as np
import torch

HDR = 512
MAGIC = 0x434D4642
E_LD = 800
BN_EPS = 1e-5


def _fold(sd, conv, bn):
    W = sd[conv + ".weight"].double()[:, :, 0, 0]
    g, b = sd[bn + ".weight"].double(), sd[bn + ".bias"].double()
    m, v = sd[bn + ".running_mean"].double(), sd[bn + ".running_var"].double()
    s = g / torch.sqrt(v + BN_EPS)
    return s[:, None] * W, b - s * m


def _padcols(W, cols):
    out = torch.zeros(W.shape[0], cols, dtype=torch.float64)
    out[:, :W.shape[1]] = W
    return out


def segments(sd, temporal=False):
    """Ordered list of (name, 2-D float64 tensor)."""
    segs = []

    def add(name, t):
        segs.append((name, t if t.dim() == 2 else t.view(1, -1)))

    for l in range(4):                                            # mse_layer (C=3)
        p = f"mse_layer.ms_ls.{l}"
        for i in range(3):
            W, t = _fold(sd, f"{p}.mlp_convs.{i}", f"{p}.mlp_bns.{i}")
            add(f"m1.{l}.W{i}", _padcols(W, 8) if i == 0 else W)
            add(f"m1.{l}.b{i}", t)
        for i in range(3):
            W, t = _fold(sd, f"{p}.mlp2_convs.{i}", f"{p}.mlp2_bns.{i}")
            add(f"m1.{l}.V{i}", W)
            add(f"m1.{l}.c{i}", t)
    W0 = sd["fc_layer.mlp_convs.0.weight"].double()[:, :, 0, 0]    # flow embedding
    add("fc.WC", W0[:, 0:256]); add("fc.WCG", W0[:, 256:512]); add("fc.WN", W0[:, 512:768]); add("fc.WNG", W0[:, 768:1024])
    add("fc.WD", _padcols(W0[:, 1024:1027], 4)); add("fc.B1", sd["fc_layer.mlp_convs.0.bias"].double())
    for i in (1, 2):
        add(f"fc.W{i+1}", sd[f"fc_layer.mlp_convs.{i}.weight"].double()[:, :, 0, 0])
        add(f"fc.B{i+1}", sd[f"fc_layer.mlp_convs.{i}.bias"].double())
    for wn in ("weightnet1", "weightnet2"):
        for i, cols in enumerate((4, 8, 8)):
            add(f"{wn}.A{i}", _padcols(sd[f"fc_layer.{wn}.mlp_convs.{i}.weight"].double()[:, :, 0, 0], cols))
            add(f"{wn}.a{i}", sd[f"fc_layer.{wn}.mlp_convs.{i}.bias"].double())
    WP = torch.zeros(2048, E_LD, dtype=torch.float64)              # set-conv #2
    WG = torch.zeros(2048, 256, dtype=torch.float64)
    T1 = torch.zeros(2048, dtype=torch.float64)
    WX = torch.zeros(2048, 4, dtype=torch.float64)
    for l in range(4):
        p = f"mse_layer2.ms_ls.{l}"
        W, t = _fold(sd, f"{p}.mlp_convs.0", f"{p}.mlp_bns.0")     # (512, 1030)
        r = slice(l * 512, (l + 1) * 512)
        WX[r, 0:3] = W[:, 0:3]
        WP[r, 768:771] = W[:, 3:6]
        WP[r, 0:256] = W[:, 6:262]
        WG[r] = W[:, 262:518]
        WP[r, 256:768] = W[:, 518:1030]
        T1[r] = t
    add("m2.WP", WP); add("m2.WG", WG); add("m2.T1", T1); add("m2.WX", WX)
    for l in range(4):
        p = f"mse_layer2.ms_ls.{l}"
        for i in (1, 2):
            W, t = _fold(sd, f"{p}.mlp_convs.{i}", f"{p}.mlp_bns.{i}")
            add(f"m2.{l}.W{i+1}", W); add(f"m2.{l}.T{i+1}", t)
        for i in range(3):
            W, t = _fold(sd, f"{p}.mlp2_convs.{i}", f"{p}.mlp2_bns.{i}")
            add(f"m2.{l}.V{i}", W); add(f"m2.{l}.c{i}", t)
    Wf, tf = _fold(sd, "fp.sf_mlp.0.0", "fp.sf_mlp.0.1")           # heads, first layer stacked [fp ; mp]
    Wm, tm = _fold(sd, "mp.sf_mlp.0.0", "mp.sf_mlp.0.1")
    add("hd.W1", torch.cat([Wf[:, 0:256], Wm[:, 0:256]], 0)); add("hd.W1G", torch.cat([Wf[:, 256:512], Wm[:, 256:512]], 0))
    add("hd.T1", torch.cat([tf, tm], 0))
    for i in (1, 2):
        for h in ("fp", "mp"):
            W, t = _fold(sd, f"{h}.sf_mlp.{i}.0", f"{h}.sf_mlp.{i}.1")
            add(f"hd.{h}.W{i+1}", W); add(f"hd.{h}.T{i+1}", t)
    add("hd.W4", torch.cat([sd["fp.conv2.weight"].double()[:, :, 0, 0], sd["mp.conv2.weight"].double()[:, :, 0, 0]], 0))
    if temporal:
        add("gru.Wih", sd["gru.weight_ih_l0"].double()); add("gru.Whh", sd["gru.weight_hh_l0"].double())
        add("gru.bih", sd["gru.bias_ih_l0"].double()); add("gru.bhh", sd["gru.bias_hh_l0"].double())
    return segs


def raflow_as_cmflow(sd):
    """RaFlow's state_dict (models/raflow.py: mse_layer, fc_layer, fd_layer.mse, fd_layer.fp) in CMFlow's key layout: the decoder's
    set-conv is CMFlow's mse_layer2, its FlowPredictor (radarflow_util.py:388-409) has FlowHead's shape, and the absent motion head
    becomes zero weights with identity BatchNorm (its scores are never read: cmf_model_forward_raflow)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("fd_layer.mse."):
            out["mse_layer2." + k[len("fd_layer.mse."):]] = v
        elif k.startswith("fd_layer.fp."):
            out["fp." + k[len("fd_layer.fp."):]] = v
        else:
            out[k] = v
    last = 512
    for i, co in enumerate((256, 128, 64)):
        out[f"mp.sf_mlp.{i}.0.weight"] = torch.zeros(co, last, 1, 1)
        out[f"mp.sf_mlp.{i}.1.weight"] = torch.ones(co); out[f"mp.sf_mlp.{i}.1.bias"] = torch.zeros(co)
        out[f"mp.sf_mlp.{i}.1.running_mean"] = torch.zeros(co); out[f"mp.sf_mlp.{i}.1.running_var"] = torch.ones(co)
        last = co
    out["mp.conv2.weight"] = torch.zeros(1, 64, 1, 1)
    return out


def pack(sd, temporal=False, raflow=False):
    """state_dict (reference key layout, any device) -> contiguous float32 numpy blob."""
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    if raflow:
        sd = raflow_as_cmflow(sd)
    segs = segments(sd, temporal)
    sizes = [(t.numel() + 3) // 4 * 4 for _, t in segs]
    blob = np.zeros(HDR + sum(sizes), dtype=np.float32)
    hdr = blob[:HDR].view(np.int32)
    assert 3 + 3 * len(segs) <= HDR
    hdr[0], hdr[1], hdr[2] = MAGIC, len(segs), int(temporal)
    off = HDR
    for i, ((_, t), sz) in enumerate(zip(segs, sizes)):
        hdr[3 + 3 * i: 6 + 3 * i] = (off, t.shape[0], t.shape[1])
        blob[off: off + t.numel()] = t.contiguous().view(-1).to(torch.float32).numpy()
        off += sz
    return blob
