"""GPU parity of the RaFlow path (models/raflow.py: the third model of models/model.py:21-27) against golden vectors produced by the
unmodified reference, in the strict-fp32 and the tensor-core builds."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from cmflow_b200._lib import check, dptr, lib, stream_ptr   # noqa: E402
from cmflow_b200.cmflow import RaFlow   # noqa: E402
from oracle import cmflow_oracle as O   # noqa: E402
from tests.helpers import case_inputs, case_weights, check_raflow_outputs, load_golden, rel_err   # noqa: E402

DEV = "cuda"


class Args:
    num_points = 256
    rigid_thres = 0.15


@pytest.mark.parametrize("precision", ["fp32", "fp16x3", "tf32x3"])
@pytest.mark.parametrize("name", ["raflow_synth_b3_n256.pt", "raflow_ckpt_b3_n256.pt"])
def test_raflow_forward_matches_reference_golden(golden_dir, name, precision):
    gold = load_golden(golden_dir, name)
    sd = case_weights(gold["meta"], golden_dir)
    if sd is None:
        pytest.skip("reference checkpoint not available")
    net = RaFlow(Args())
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV)
    net.set_precision(precision)
    pc1, pc2, ft1, ft2 = (t.to(DEV) for t in case_inputs(gold["meta"])[:4])
    with torch.no_grad():
        output, sf_agg, pre_trans, mask_s = net(pc1, pc2, ft1, ft2, gold["interval"].to(DEV))
    B, N = gold["meta"]["B"], gold["meta"]["N"]
    assert output.shape == (B, 3, N) and sf_agg.shape == (B, 3, N) and pre_trans.shape == (B, 4, 4) and mask_s.dtype == torch.bool
    out = {"output": output.cpu(), "sf_agg": sf_agg.cpu(), "pre_trans": pre_trans.cpu(), "mask_s": mask_s.cpu()}
    E = net.tap("E", (B, N, 800)).cpu()
    assert rel_err(E[0, ::4, 0:256].t(), gold["f1_sub"], per_pair=False) <= 1e-4
    assert rel_err(E[0, ::4, 256:768].t(), gold["cor_sub"], per_pair=False) <= 1e-4
    assert rel_err(net.tap("prop", (B, N, 256)).cpu()[0, ::4].t(), gold["prop_sub"], per_pair=False) <= 1e-4
    print(name, precision, check_raflow_outputs(out, gold))


def test_raflow_refine_operator_matches_oracle():
    """cmf_raflow_refine alone (raflow.py:79-156) on random flows, both branches, ragged N, a zero radial velocity (division by zero -> no inlier)."""
    g = torch.Generator().manual_seed(7)
    B, N = 6, 203
    pc1 = torch.rand(B, 3, N, generator=g) * torch.tensor([50.0, 40.0, 4.0]).view(1, 3, 1) - torch.tensor([0.0, 20.0, 2.0]).view(1, 3, 1)
    flow = torch.randn(B, 3, N, generator=g) * 0.05 + torch.tensor([0.3, 0.02, 0.0]).view(1, 3, 1)
    ft1 = torch.randn(B, 3, N, generator=g) * 2.0
    ft1[0, 0, 5] = 0.0
    interval = torch.tensor([0.1, 0.1, 0.1, 0.6, 0.1, 0.05])
    # radial velocities consistent with the flow for half of the pairs -> many inliers there, few elsewhere
    proj = (flow * pc1).sum(1) / pc1.norm(dim=1)
    ft1[:3, 0] = proj[:3] / interval[:3].view(3, 1) * (1 + 0.05 * torch.randn(3, N, generator=g))
    ft1[0, 0, 5] = 0.0
    ref = {}
    pcd = pc1
    trans = O.raflow_rigid_transform(pcd, pcd + flow, torch.ones(B, N))
    sf_rg = O.rigid_to_flow(pcd, trans)
    ratio = ((ft1[:, 0] * interval.unsqueeze(1) - (sf_rg * pcd).sum(1) / pcd.norm(dim=1)) / ft1[:, 0]).abs()
    safe = (ratio - 0.15).abs() > 1e-4
    mask_ref = ratio < 0.15
    d = {k: v.to(DEV).contiguous() for k, v in dict(pc1=pc1, ft1=ft1, flow=flow, interval=interval).items()}
    sf = torch.empty(B, 3, N, device=DEV); T = torch.empty(B, 4, 4, device=DEV); mask = torch.empty(B, N, dtype=torch.uint8, device=DEV)
    check(lib().cmf_raflow_refine(B, N, dptr(d["pc1"]), dptr(d["ft1"]), dptr(d["flow"]), dptr(d["interval"]), 0.15, 0.25,
                                  dptr(sf), dptr(T), dptr(mask), stream_ptr()))
    torch.cuda.synchronize()
    mask = mask.cpu().bool()
    assert torch.equal(mask | ~safe, mask_ref | ~safe)
    assert not mask[0, 5]
    frac = mask.float().mean(1)
    assert (frac > 0.25).any() and (frac < 0.25).any()
    # the rest of the module from the GPU's own mask (so that a legitimately flipped near-threshold bit does not move the fit)
    for b in range(B):
        if frac[b] > 0.25:
            Tb = O.raflow_rigid_transform(pcd[b:b + 1].double(), (pcd + flow)[b:b + 1].double(), mask[b:b + 1])[0].float()
            want = O.rigid_to_flow(pcd[b:b + 1], Tb.unsqueeze(0))[0]
            want[:, ~mask[b]] = flow[b][:, ~mask[b]]
        else:
            Tb = O.raflow_rigid_transform(pcd[b:b + 1].double(), (pcd + flow)[b:b + 1].double(), torch.ones(1, N))[0].float()
            want = flow[b]
        assert rel_err(T[b, :3].cpu().unsqueeze(0), Tb[:3].unsqueeze(0)) <= 1e-4
        assert rel_err(sf[b].cpu().unsqueeze(0), want.unsqueeze(0)) <= 1e-4
