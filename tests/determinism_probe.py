"""Diagnostic (not a test): where does a fp16x3 forward stop being bit-reproducible?

    python tests/determinism_probe.py

Runs the batch-independence case of tests/test_gpu_tc_gemm.py (6 pairs vs the first 2 of them) and the same batch twice, and
prints, per engine stage tap, how many elements differ and by how much."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cmflow_b200.cmflow import CMFlow  # noqa: E402
from cmflow_b200.synth import make_pairs, synthetic_state_dict  # noqa: E402


class Args:
    num_points = 256
    stat_thres = 0.5


TAPS = (("u1", 512), ("u2", 512), ("cost1", 512), ("E", 800), ("f2", 256), ("P", 2048), ("prop", 256), ("hd3", 128), ("flow", 3))


def forward(net, inp, B, N):
    with torch.no_grad():
        out = net(*[t.cuda() for t in inp], None, "test")
    torch.cuda.synchronize()
    taps = {k: net.tap(k, (B * N, c)).cpu() for k, c in TAPS}
    taps["sf"] = out[0].cpu()
    taps["T"] = out[2].cpu()
    return taps


def diff(a, b, rows=None):
    res = []
    for k in a:
        x, y = a[k], b[k]
        if rows is not None and k not in ("sf", "T"):
            x, y = x[:rows], y[:rows]
        elif rows is not None:
            n = min(x.shape[0], y.shape[0]); x, y = x[:n], y[:n]
        if k == "flow" and rows is not None:
            continue                                     # planar (B,3,N): rows do not line up across batch sizes
        ne = (x != y).sum().item()
        res.append(f"{k}:{ne}" + (f"({(x - y).abs().max().item():.1e})" if ne else ""))
    return " ".join(res)


def main():
    net = CMFlow(Args()); net.load_state_dict(synthetic_state_dict(0)); net = net.to("cuda:0")
    net.set_precision("fp16x3")
    N = 256
    inp = make_pairs(6, N, seed=11)
    big = [t.clone() for t in inp[:4]]
    big[0][3:] *= 3.0; big[1][3:] *= 3.0
    w1 = forward(net, big, 6, N)
    w2 = forward(net, big, 6, N)
    print("run-to-run  (6 pairs):", diff(w1, w2))
    s1 = forward(net, [t[:2] for t in big], 2, N)
    s2 = forward(net, [t[:2] for t in big], 2, N)
    print("run-to-run  (2 pairs):", diff(s1, s2))
    print("2 vs 6 pairs         :", diff(s1, w1, rows=2 * N))
    for _ in range(3):
        w3 = forward(net, big, 6, N)
        print("again       (6 pairs):", diff(w1, w3))


if __name__ == "__main__":
    main()
