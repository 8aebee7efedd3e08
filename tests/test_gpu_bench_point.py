"""GPU parity AT the bench configurations, in the bench precision (fp16x3 = tcgen05 3xFP16 split), against the CPU oracle:

  BASELINE.json configs[1]  CMFlow, N=256, batch=256/GPU      -- sampled pairs spread over chunks and cluster tiles
  configs[3]               CMFlow-T, 3-frame clips, 64 clips/GPU (= batch 512 over 8 GPUs)
  configs[4]               N=4096, batch=64

The small golden cases give every CTA cluster at most one tile; these sizes run the persistent multi-tile rings of the tensor-core
kernels (74 clusters x many tiles), which is where a ring-protocol bug would show.  Bars: neighbour sets identical, flow / transform /
scores <= 1e-4 (tests/helpers.py), and bit-equality with a small-batch run of the same pairs in the same precision (pairs are independent).
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from cmflow_b200.cmflow import CMFlow, CMFlow_T   # noqa: E402
from cmflow_b200.synth import make_pairs, synthetic_state_dict   # noqa: E402
from oracle import cmflow_oracle as O   # noqa: E402
from tests.helpers import check_outputs, knn_sets_equal   # noqa: E402

DEV = "cuda"


class Args:
    num_points = 256
    stat_thres = 0.5


def run(net, inp, g=None):
    pc1, pc2, ft1, ft2 = (t.to(DEV) for t in inp[:4])
    with torch.no_grad():
        if isinstance(net, CMFlow_T):
            sf, cls, T, mask, g = net(pc1, pc2, ft1, ft2, None, "test", g)
        else:
            sf, cls, T, mask = net(pc1, pc2, ft1, ft2, None, "test")
    return {"sf_agg": sf.cpu(), "stat_cls": cls.cpu(), "pre_trans": T.cpu(), "mask": mask.cpu(), "gfeat": g}


def take(out, sel):
    return {k: (v[sel] if torch.is_tensor(v) else v) for k, v in out.items()}


def test_batch256_fp16x3_sampled_pairs_match_oracle():
    """configs[1] in the precision bench.py times."""
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    sd = synthetic_state_dict(0)
    net = CMFlow(Args()); net.load_state_dict(sd); net = net.to(DEV); net.set_precision("fp16x3")
    B, N = 256, 256
    inp = make_pairs(B, N, seed=1234)                       # bench.py's first input set
    out = run(net, inp)
    knn12 = net.tap("knn12", (B, N, 8), torch.int32).cpu()
    knn11 = net.tap("knn11", (B, N, 8), torch.int32).cpu()
    assert net.workspace_bytes() <= 10 * (1 << 30), net.workspace_bytes()            # tensor-core modes skip the fp32-only buffers
    sel = torch.tensor([0, 37, 73, 101, 128, 170, 203, 255])
    ref = O.cmflow_forward(sd, *(t[sel] for t in inp[:4]), return_intermediates=True)
    assert knn_sets_equal(knn12[sel], ref["knn12"].long().sort(-1)[0])
    assert knn_sets_equal(knn11[sel], ref["knn11"].long().sort(-1)[0])
    errs = check_outputs(take(out, sel), ref)
    print("B=256 fp16x3 sampled pairs vs oracle:", errs)
    # batch independence, bitwise, in the same precision: the first two pairs alone, and the last two alone
    for lo in (0, B - 2):
        small = run(net, tuple(t[lo:lo + 2].contiguous() for t in inp[:4]))
        for k in ("sf_agg", "stat_cls", "pre_trans", "mask"):
            assert torch.equal(small[k], out[k][lo:lo + 2]), (k, lo)
    # strict-fp32 build of the same batch agrees with the tensor-core build far inside the bar
    net.set_precision("fp32")
    out32 = run(net, inp)
    d = (out32["sf_agg"] - out["sf_agg"]).abs().amax((1, 2)) / out32["sf_agg"].abs().amax((1, 2))
    safe = ((out32["stat_cls"] - 0.5).abs() > 1e-4).all(2).squeeze(1)               # pairs without a point on the mask threshold
    print("fp16x3 vs fp32 build, max rel flow difference over the 256 pairs:", d[safe].max().item())
    assert d[safe].max() <= 2e-4           # each is within 1e-4 of the oracle


def test_temporal_64_clips_fp16x3_match_oracle():
    """configs[3] per-GPU share: 64 three-frame clips, GRU state carried (clip_util.py:218-233)."""
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    sd = synthetic_state_dict(3, temporal=True)
    net = CMFlow_T(Args()); net.load_state_dict(sd); net = net.to(DEV); net.set_precision("fp16x3")
    B, N = 64, 256
    frames = [make_pairs(B, N, seed=500 + f) for f in range(3)]
    sel = torch.tensor([0, 21, 42, 63])
    g, gref = None, None
    for f, inp in enumerate(frames):
        out = run(net, inp, g)
        ref = O.cmflow_forward(sd, *(t[sel] for t in inp[:4]), temporal=True, gfeat_prev=gref)
        errs = check_outputs(take(out, sel), ref)
        ge = (out["gfeat"].cpu()[sel] - ref["gfeat"]).abs().max().item() / ref["gfeat"].abs().max().item()
        print(f"clip frame {f}:", errs, "gfeat", ge)
        assert ge <= 1e-4
        g, gref = out["gfeat"], ref["gfeat"]


def test_dense_batch64_fp16x3():
    """configs[4]: N=4096, batch=64.  One pair against the oracle (unfused: 0.5 GB of grouped tensors per pair on the CPU), the others through
    bit-equality with single-pair runs."""
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    sd = synthetic_state_dict(0)
    net = CMFlow(Args()); net.load_state_dict(sd); net = net.to(DEV); net.set_precision("fp16x3")
    B, N = 64, 4096
    inp = make_pairs(B, N, seed=11, dense=True)
    out = run(net, inp)
    assert all(torch.isfinite(out[k]).all() for k in ("sf_agg", "stat_cls", "pre_trans"))
    for i in (0, 31, 63):
        one = run(net, tuple(t[i:i + 1].contiguous() for t in inp[:4]))
        for k in ("sf_agg", "stat_cls", "pre_trans", "mask"):
            assert torch.equal(one[k], out[k][i:i + 1]), (k, i)
    ref = O.cmflow_forward(sd, *(t[63:64] for t in inp[:4]))
    print("N=4096 B=64 pair 63 vs oracle:", check_outputs(take(out, slice(63, 64)), ref))
