"""tcgen05 3xTF32 GEMM (csrc/tc_gemm.cu) against an fp64 reference, and the engine in tf32x3 mode against the
golden vectors / the strict-fp32 engine."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from cmflow_b200._lib import check, dptr, lib, stream_ptr   # noqa: E402
from cmflow_b200.cmflow import CMFlow, CMFlow_T   # noqa: E402
from tests.helpers import case_inputs, case_weights, check_outputs, load_golden, rel_err   # noqa: E402

DEV = "cuda"


def tc_gemm(W, X, bias, act, ldx=None):
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
    check(lib().cmf_test_tc_gemm(M, K, cols, dptr(Wd), K, dptr(Xp), ldx, dptr(bd), act, dptr(out), M, dptr(scratch), stream_ptr()))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,K,cols", [(128, 32, 256), (128, 64, 256), (256, 512, 1000), (512, 256, 70000), (64, 256, 300),
                                      (2048, 771, 513), (128, 32, 1)])
def test_tc_gemm_matches_fp64(M, K, cols):
    g = torch.Generator().manual_seed(M * 7 + K)
    W = torch.randn(M, K, generator=g) / K ** 0.5
    X = torch.randn(cols, K, generator=g) * 3
    bias = torch.randn(M, generator=g)
    got = tc_gemm(W, X.to(DEV), bias, 1).cpu().double()
    want = torch.relu(X.double() @ W.double().t() + bias.double())
    err = (got - want).abs().max().item() / want.abs().max().item()
    print(M, K, cols, "max rel err", err)
    assert not torch.isnan(got).any()
    # fp32-class accuracy: single-pass TF32 would be ~5e-4.  The residual (~4e-6 at K=512) is the tensor core's
    # fp32 accumulator, which truncates instead of rounding on every K=8 step (3 x K/8 accumulations per output).
    assert err < 1e-5, err


@pytest.mark.parametrize("M,cols", [(128, 256), (256, 256), (256, 700), (512, 1300)])
def test_tc_gemm_identity_layout(M, cols):
    """W = I picks X apart element by element: catches any swizzle / descriptor / lane-mapping slip exactly.
    M = 128 runs the one-CTA kernel, M % 256 == 0 the CTA-pair (cta_group::2) kernel."""
    K = M
    X = torch.arange(cols * K, dtype=torch.float32).reshape(cols, K) / 64.0     # exactly representable in TF32 hi+lo
    got = tc_gemm(torch.eye(M), X.to(DEV), None, 0).cpu()
    assert torch.equal(got, X)


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


@pytest.mark.parametrize("name", ["cmflow_synth_b2_n256.pt", "cmflow_synth_w1_b2_n256.pt", "cmflow_synth_b3_n200.pt",
                                  "cmflow_synth_b2_n40.pt", "cmflow_ckpt_b2_n256.pt"])
def test_tf32x3_forward_matches_reference_golden(golden_dir, name):
    gold = load_golden(golden_dir, name)
    sd = case_weights(gold["meta"], golden_dir)
    if sd is None:
        pytest.skip("reference checkpoint not available")
    net = CMFlow(Args()); net.load_state_dict(sd); net = net.to(DEV)
    net.set_precision("tf32x3")
    inp = case_inputs(gold["meta"])
    out = run(net, inp)
    errs = check_outputs(out, gold)
    print(name, "tf32x3", errs)
    net32 = CMFlow(Args()); net32.load_state_dict(sd); net32 = net32.to(DEV)
    ref = run(net32, inp)
    print("vs strict fp32 engine: flow", rel_err(out["sf_agg"], ref["sf_agg"]), "trans", rel_err(out["pre_trans"][:, :3], ref["pre_trans"][:, :3]))


def test_tf32x3_temporal_matches_reference_golden(golden_dir):
    gold = load_golden(golden_dir, "cmflow_t_synth_b2_n256.pt")
    sd = case_weights(gold["meta"], golden_dir)
    net = CMFlow_T(Args()); net.load_state_dict(sd); net = net.to(DEV)
    net.set_precision("tf32x3")
    inp = case_inputs(gold["meta"])
    g = None
    for step in gold["steps"]:
        out = run(net, inp, g)
        check_outputs(out, step)
        g = out["gfeat"]
