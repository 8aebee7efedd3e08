"""Numerical study (CPU emulation, not a test): how much end-to-end error do cheaper operand splits of the tensor-core
GEMMs cost?  VERDICT r01 item 4: "measure the FP8-correction split instead of dismissing it".

    python tests/split_precision_study.py [--ckpt] [--layers all|sc2l2]

The engine's fp16x3 mode computes every big 1x1 conv as  Ah.Bh + Ah.Bl + Al.Bh  (A, B split into fp16 hi + fp16 lo after exact
power-of-two scaling, fp32 accumulation).  Variants emulated here on the oracle's direct-form forward, all against the fp64
oracle as truth:

  fp16x3     : the shipped split (three kind::f16 MMAs per product)
  fp16+2xfp8 : hi x hi in fp16, both cross terms with e4m3 copies of BOTH factors (kind::f8f6f4 at twice the rate: 2 tensor
               units per MAC instead of 3).  hi8 = e4m3(hi * 2^-6), lo8 = e4m3(lo * 2^+6) so that the product keeps the
               accumulator's scale and both stay inside e4m3's range (|x| <= 448).
  fp16x2w    : weights split (hi + lo), activations single fp16 (two MMAs)
  fp16x1     : plain fp16 operands (one MMA)

Emulation notes: products of fp16 (or e4m3) values are exact in fp32/fp64 and the tensor core accumulates in fp32; the
accumulation-order error (~1e-7 relative) is below what is studied here, so the emulated sums are taken in fp64.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cmflow_b200.synth import make_pairs, synthetic_state_dict  # noqa: E402
from oracle import cmflow_oracle as O  # noqa: E402

TARGET_EXP = 13          # scaled max magnitude in [2^13, 2^14)


def pow2_scale(amax, slack_bits=0):
    """power of two s with amax * s in [2^(13 - slack), 2^(14 - slack))"""
    e = torch.floor(torch.log2(amax.clamp_min(1e-30)))
    return torch.pow(2.0, (TARGET_EXP - slack_bits) - e)


def q16(x):
    return x.to(torch.float16).to(torch.float64)


def q8(x):
    return x.clamp(-448.0, 448.0).to(torch.float32).to(torch.float8_e4m3fn).to(torch.float64)


def split_matmul(W, X, variant, slack_bits):
    """W (O,C) fp64, X (B,C,M) fp64 -> (B,O,M): emulated split-precision product with per-row weight scales and a per-pair
    activation scale (the engine's bound-based scale can sit a few binades below the true maximum: slack_bits)."""
    sw = pow2_scale(W.abs().amax(1, keepdim=True))                      # (O,1)
    sx = pow2_scale(X.abs().amax((1, 2), keepdim=True), slack_bits)     # (B,1,1)
    Ws, Xs = W * sw, X * sx
    Wh = q16(Ws); Wl = q16(Ws - Wh)
    Xh = q16(Xs); Xl = q16(Xs - Xh)
    mm = lambda a, b: torch.einsum("oc,bcm->bom", a, b)
    if variant == "fp16x3":
        acc = mm(Wh, Xh) + mm(Wh, Xl) + mm(Wl, Xh)
    elif variant == "fp16+2xfp8":
        acc = mm(Wh, Xh) + mm(q8(Wh * 2.0 ** -6), q8(Xl * 2.0 ** 6)) + mm(q8(Wl * 2.0 ** 6), q8(Xh * 2.0 ** -6))
    elif variant == "fp16x2w":
        acc = mm(Wh, Xh) + mm(Wl, Xh)
    elif variant == "fp16x1":
        acc = mm(Wh, Xh)
    else:
        raise ValueError(variant)
    return acc / sw.view(1, -1, 1) / sx


def make_conv(variant, which, slack_bits, base_variant="fp16x3"):
    def conv(sd, name, x, dtype):
        w = sd[name + ".weight"].to(dtype)[:, :, 0, 0]
        big = w.shape[1] >= 128 and w.shape[0] >= 64
        if not big:
            y = torch.einsum("oc,bc...->bo...", w, x)
        else:
            v = variant if (which == "all" or (which == "sc2l2" and ".mlp_convs.1" in name and "mse_layer2" in name)) else base_variant
            sh = x.shape
            y = split_matmul(w.double(), x.reshape(sh[0], sh[1], -1).double(), v, slack_bits).reshape(sh[0], w.shape[0], *sh[2:]).to(dtype)
        b = sd.get(name + ".bias")
        if b is not None:
            y = y + b.to(dtype).view(1, -1, *([1] * (x.dim() - 2)))
        return y
    return conv


def rel(a, b):
    scale = b.abs().flatten(1).max(1)[0].view(-1, *([1] * (b.dim() - 1)))
    return ((a - b).abs() / scale).max().item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt", action="store_true", help="reference's cmflow_cvpr weights (needs /root/reference) instead of seeded synthetic ones")
    ap.add_argument("--pairs", type=int, default=2)
    ap.add_argument("--points", type=int, default=256)
    ap.add_argument("--seeds", type=int, default=3)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 8)
    if args.ckpt:
        sd = torch.load("/root/reference/checkpoints/cmflow_cvpr/models/model.best.t7", map_location="cpu", weights_only=True)
    else:
        sd = synthetic_state_dict(0)
    orig = O._conv
    print(f"weights={'cmflow_cvpr' if args.ckpt else 'synthetic'} pairs={args.pairs} N={args.points}")
    print(f"{'variant':<14}{'layers':<8}{'slack':<6}{'seed':<6}{'flow':>10}{'trans':>10}{'cls':>10}")
    for seed in range(args.seeds):
        pc1, pc2, ft1, ft2, _ = make_pairs(args.pairs, args.points, seed=1234 + seed)
        with torch.no_grad():
            truth = O.cmflow_forward(sd, pc1, pc2, ft1, ft2, dtype=torch.float64)
            safe = ((truth["stat_cls"] - 0.5).abs() > 1e-3).squeeze(1)
            ref32 = O.cmflow_forward(sd, pc1, pc2, ft1, ft2, dtype=torch.float32)

            def errs(out):
                a = torch.where(safe.unsqueeze(1), out["sf_agg"].double(), truth["sf_agg"])
                return rel(a, truth["sf_agg"]), rel(out["pre_trans"][:, :3].double(), truth["pre_trans"][:, :3]), \
                    (out["stat_cls"].double() - truth["stat_cls"]).abs().max().item()
            e = errs(ref32)
            print(f"{'torch fp32':<14}{'-':<8}{'-':<6}{seed:<6}{e[0]:>10.2e}{e[1]:>10.2e}{e[2]:>10.2e}")
            for variant, which, slack in (("fp16x3", "all", 0), ("fp16x3", "all", 3), ("fp16+2xfp8", "sc2l2", 0), ("fp16+2xfp8", "sc2l2", 3),
                                          ("fp16+2xfp8", "all", 0), ("fp16+2xfp8", "all", 3), ("fp16x2w", "all", 0), ("fp16x1", "all", 0)):
                O._conv = make_conv(variant, which, slack)
                try:
                    out = O.cmflow_forward(sd, pc1, pc2, ft1, ft2, dtype=torch.float64)
                finally:
                    O._conv = orig
                e = errs(out)
                print(f"{variant:<14}{which:<8}{slack:<6}{seed:<6}{e[0]:>10.2e}{e[1]:>10.2e}{e[2]:>10.2e}", flush=True)


if __name__ == "__main__":
    main()
