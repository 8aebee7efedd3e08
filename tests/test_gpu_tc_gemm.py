"""tcgen05 3xTF32 GEMM (csrc/tc_gemm.cu) against an fp64 reference, and the engine in tf32x3 mode against the
golden vectors / the strict-fp32 engine."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from cmflow_b200._lib import check, dptr, lib, stream_ptr   # noqa: E402
from cmflow_b200.cmflow import CMFlow, CMFlow_T   # noqa: E402
from tests.helpers import case_inputs, case_weights, check_outputs, load_golden, rel_err   # noqa: E402

DEV = "cuda"


def tc_gemm(W, X, bias, act, ldx=None, fmt=0, cols_per_pair=0, amax_in=None, amax_out=None):
    M, K = W.shape
    cols = X.shape[0]
    kp = (K + 31) // 32 * 32
    ldx = ldx or kp
    Xp = torch.zeros(cols, ldx, device=DEV)
    Xp[:, :K] = X
    out = torch.full((cols, M), float("nan"), device=DEV)
    scratch = torch.empty(lib().cmf_test_tc_tiled_floats(M, K), device=DEV)
    Wd = W.to(DEV).contiguous()
    bd = bias.to(DEV) if bias is not None else None
    check(lib().cmf_test_tc_gemm_fmt(fmt, M, K, cols, dptr(Wd), K, dptr(Xp), ldx, dptr(bd), act, dptr(out), M, dptr(scratch),
                                     cols_per_pair, dptr(amax_in), dptr(amax_out), stream_ptr()))
    torch.cuda.synchronize()
    return out


FMTS = [0, 1]      # 0 = 3xTF32 (kind::tf32), 1 = 3xFP16 (kind::f16, power-of-two scaled operands)


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("M,K,cols", [(128, 32, 256), (128, 64, 256), (256, 512, 1000), (512, 256, 70000), (64, 256, 300),
                                      (2048, 771, 513), (128, 32, 1)])
def test_tc_gemm_matches_fp64(M, K, cols, fmt):
    g = torch.Generator().manual_seed(M * 7 + K)
    W = torch.randn(M, K, generator=g) / K ** 0.5
    X = torch.randn(cols, K, generator=g) * 3
    bias = torch.randn(M, generator=g)
    got = tc_gemm(W, X.to(DEV), bias, 1, fmt=fmt).cpu().double()
    want = torch.relu(X.double() @ W.double().t() + bias.double())
    err = (got - want).abs().max().item() / want.abs().max().item()
    print(M, K, cols, "fmt", fmt, "max rel err", err)
    assert not torch.isnan(got).any()
    # fp32-class accuracy: single-pass TF32 / FP16 would be ~5e-4.  The residual (~4e-6 at K=512) is the tensor core's
    # fp32 accumulator, which truncates instead of rounding on every K step (3 x K/8 or 3 x K/16 accumulations per output).
    assert err < 1e-5, err


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("M,cols", [(128, 256), (256, 256), (256, 700), (512, 1300)])
def test_tc_gemm_identity_layout(M, cols, fmt):
    """W = I picks X apart element by element: catches any swizzle / descriptor / lane-mapping slip exactly.
    M = 128 runs the one-CTA kernel, M % 256 == 0 the CTA-pair (cta_group::2) kernel."""
    K = M
    X = (torch.arange(cols * K, dtype=torch.float32).reshape(cols, K) % 8191) / 64.0     # exactly representable as hi+lo in either format
    got = tc_gemm(torch.eye(M), X.to(DEV), None, 0, fmt=fmt).cpu()
    assert torch.equal(got, X)


@pytest.mark.parametrize("M,K,cols,cpp", [(256, 512, 1024, 256), (512, 256, 1000, 200), (128, 96, 700, 100), (64, 256, 600, 96)])
def test_fp16x3_dynamic_range_and_pair_maxima(M, K, cols, cpp):
    """fp16 overflows at 65504 and loses precision below 6e-5: groups of `cpp` activation rows ('frame pairs') with magnitudes
    from 1e-6 to 1e+7 must come out with the same relative accuracy thanks to the per-group power-of-two scale derived from the
    measured maxima; weight rows spanning 1e-5..1e+4 likewise (per-row scale).  The epilogue's per-group max|out| is exact."""
    g = torch.Generator().manual_seed(M + K + cols)
    ng = (cols + cpp - 1) // cpp
    mag = 10.0 ** torch.linspace(-6, 7, ng)
    X = torch.randn(cols, K, generator=g) * mag.repeat_interleave(cpp)[:cols, None]
    wmag = 10.0 ** torch.linspace(-5, 4, M)
    W = torch.randn(M, K, generator=g) * wmag[:, None] / K ** 0.5
    amax_in = torch.stack([X[i * cpp:(i + 1) * cpp].abs().max() for i in range(ng)]).to(DEV)
    amax_out = torch.zeros(ng, dtype=torch.int32, device=DEV)
    got = tc_gemm(W, X.to(DEV), None, 0, fmt=1, cols_per_pair=cpp, amax_in=amax_in, amax_out=amax_out).cpu().double()
    want = X.double() @ W.double().t()
    assert torch.isfinite(got).all()
    # error relative to the natural scale of each output: |x|_2 of the row times |w|_2 of the weight row
    scale = X.double().norm(dim=1)[:, None] * W.double().norm(dim=1)[None, :]
    err = ((got - want).abs() / scale).max().item()
    print(M, K, cols, cpp, "max err / (|x||w|)", err)
    assert err < 2e-6, err
    got_max = amax_out.view(torch.float32).cpu()
    want_max = torch.stack([got[i * cpp:(i + 1) * cpp].abs().max() for i in range(ng)]).float()
    assert torch.equal(got_max, want_max)


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


@pytest.mark.parametrize("precision", ["tf32x3", "fp16x3"])
@pytest.mark.parametrize("name", ["cmflow_synth_b2_n256.pt", "cmflow_synth_w1_b2_n256.pt", "cmflow_synth_b3_n200.pt",
                                  "cmflow_synth_b2_n40.pt", "cmflow_ckpt_b2_n256.pt"])
def test_split_precision_forward_matches_reference_golden(golden_dir, name, precision):
    gold = load_golden(golden_dir, name)
    sd = case_weights(gold["meta"], golden_dir)
    if sd is None:
        pytest.skip("reference checkpoint not available")
    net = CMFlow(Args()); net.load_state_dict(sd); net = net.to(DEV)
    net.set_precision(precision)
    inp = case_inputs(gold["meta"])
    out = run(net, inp)
    errs = check_outputs(out, gold)
    print(name, precision, errs)
    net32 = CMFlow(Args()); net32.load_state_dict(sd); net32 = net32.to(DEV)
    ref = run(net32, inp)
    print("vs strict fp32 engine: flow", rel_err(out["sf_agg"], ref["sf_agg"]), "trans", rel_err(out["pre_trans"][:, :3], ref["pre_trans"][:, :3]))


@pytest.mark.parametrize("precision", ["tf32x3", "fp16x3"])
def test_split_precision_temporal_matches_reference_golden(golden_dir, precision):
    gold = load_golden(golden_dir, "cmflow_t_synth_b2_n256.pt")
    sd = case_weights(gold["meta"], golden_dir)
    net = CMFlow_T(Args()); net.load_state_dict(sd); net = net.to(DEV)
    net.set_precision(precision)
    inp = case_inputs(gold["meta"])
    g = None
    for step in gold["steps"]:
        out = run(net, inp, g)
        check_outputs(out, step)
        g = out["gfeat"]


@pytest.mark.parametrize("precision", ["tf32x3", "fp16x3"])
def test_split_precision_stage_taps_match_fp64_emulation(golden_dir, precision):
    """Stage boundaries of the tensor-core pipeline vs the fp64 replay of the same packed weights."""
    from cmflow_b200 import weights
    from tests.pipeline_emulator import emulate
    gold = load_golden(golden_dir, "cmflow_synth_b2_n256.pt")
    sd = case_weights(gold["meta"], golden_dir)
    net = CMFlow(Args()); net.load_state_dict(sd); net = net.to(DEV)
    net.set_precision(precision)
    inp = case_inputs(gold["meta"])
    out = run(net, inp)
    B, N = 2, 256
    em = emulate(weights.pack(sd, False), *inp[:4], dtype=torch.float64)
    E = net.tap("E", (B, N, 800)).cpu()
    for name, got, want in (("f1", E[..., 0:256], em["f1"]), ("f2", net.tap("f2", (B, N, 256)).cpu(), em["f2"]),
                            ("cor", E[..., 256:768], em["cor"]), ("prop", net.tap("prop", (B, N, 256)).cpu(), em["prop"]),
                            ("flow", net.tap("flow", (B, 3, N)).cpu(), em["flow"])):
        e = rel_err(got, want)
        print(precision, name, e)
        assert e <= 5e-5, (name, e)


def test_fp16x3_is_batch_independent():
    """Operand scales are chosen per frame pair, so a pair's result does not depend on what else is in the batch."""
    from cmflow_b200.synth import make_pairs, synthetic_state_dict
    net = CMFlow(Args()); net.load_state_dict(synthetic_state_dict(0)); net = net.to(DEV)
    net.set_precision("fp16x3")
    inp = make_pairs(6, 256, seed=11)
    big = [t.clone() for t in inp[:4]]
    big[0][3:] *= 3.0; big[1][3:] *= 3.0                 # other pairs in the batch with very different magnitudes
    whole = run(net, big)
    small = run(net, tuple(t[:2] for t in big))
    for k in ("sf_agg", "stat_cls", "pre_trans", "mask"):
        assert torch.equal(small[k], whole[k][:2]), k


@pytest.mark.parametrize("pairs", [6, 96])
def test_fp16x3_is_run_to_run_reproducible(pairs):
    """Every tensor-core stage accumulates in one fixed order (a single MMA-issuing thread per CTA pair): the same batch twice gives the same
    bits, stage by stage.  (Two alternating issuer warps with a fast, elect-predicated issue did not: tests/determinism_probe.py.)"""
    from cmflow_b200.synth import make_pairs, synthetic_state_dict
    net = CMFlow(Args()); net.load_state_dict(synthetic_state_dict(0)); net = net.to(DEV)
    net.set_precision("fp16x3")
    N = 256
    inp = make_pairs(pairs, N, seed=5)
    taps = (("u1", 512), ("u2", 512), ("cost1", 512), ("P", 2048), ("prop", 256), ("hd3", 128))

    def once():
        out = run(net, inp)
        out.update({k: net.tap(k, (pairs * N, c)).cpu() for k, c in taps})
        return out
    a = once()
    for _ in range(2):
        b = once()
        for k in ("sf_agg", "stat_cls", "pre_trans", "mask") + tuple(k for k, _ in taps):
            assert torch.equal(a[k], b[k]), k
