"""The self-supervised losses (cmflow_b200/losses.py) against the UNMODIFIED reference losses/radar_loss.py running on the same GPU
(staged reference Python over its own kernels, oracle/ref_model.py): values and gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from cmflow_b200 import losses as L   # noqa: E402
from cmflow_b200.synth import make_pairs   # noqa: E402
from oracle import ref_model as RM   # noqa: E402

DEV = "cuda"


def ref_losses():
    if not RM.available("cuda"):
        pytest.skip("oracle/_ref (reference kernels + staged reference Python) not built")
    RM.strict_fp32()
    RM.load("cuda")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from losses import radar_loss
    return radar_loss


def inputs(B, N, M, seed):
    a, b = make_pairs(B, N, seed=seed), make_pairs(B, M, seed=seed + 1)
    g = torch.Generator().manual_seed(seed)
    pc1, pc2 = a[0].to(DEV), b[1].to(DEV)
    flow = (torch.randn(B, 3, N, generator=g) * 0.3).to(DEV)
    vel = a[2][:, 0].to(DEV)
    return pc1, pc2, flow, vel


@pytest.mark.parametrize("B,N,M", [(4, 256, 256), (2, 200, 173), (1, 64, 64)])
def test_losses_match_reference_values_and_gradients(B, N, M):
    R = ref_losses()
    pc1, pc2, flow, vel = inputs(B, N, M, seed=31 + N)
    # the chamfer loss only counts points whose kernel density exceeds zeta: shrink the clouds so that most points are inliers
    d1, d2 = (pc1 * 0.12).contiguous(), (pc2 * 0.12).contiguous()
    for name, mine, theirs, args in (
            ("chamfer", L.SoftChamferLoss(), R.SoftChamferLoss(), lambda f: (d1, d2, d1 + f)),
            ("smoothness", L.SpatialSmoothnessLoss(), R.SpatialSmoothnessLoss(), lambda f: (pc1, f)),
            ("radial", L.RadialDisplacementLoss(), R.RadialDisplacementLoss(), lambda f: (pc1, f, vel))):
        f1 = flow.clone().requires_grad_(True)
        f2 = flow.clone().requires_grad_(True)
        v1 = mine(*args(f1))
        v2 = theirs(*args(f2))
        v1.backward(); v2.backward()
        rel = abs(v1.item() - v2.item()) / max(abs(v2.item()), 1e-12)
        gerr = (f1.grad - f2.grad).abs().max().item() / max(f2.grad.abs().max().item(), 1e-12)
        print(name, (B, N, M), "value", v1.item(), v2.item(), "rel", rel, "grad rel", gerr)
        assert v2.item() > 0 and f2.grad.abs().max() > 0, name          # the case exercises the loss
        assert rel <= 1e-4 and gerr <= 1e-4, name


def test_self_supervised_sum_matches_reference():
    R = ref_losses()
    pc1, pc2, flow, vel = inputs(3, 256, 256, seed=5)
    pc1, pc2 = (pc1 * 0.12).contiguous(), (pc2 * 0.12).contiguous()
    f1 = flow.clone().requires_grad_(True)
    f2 = flow.clone().requires_grad_(True)
    t1, items1 = L.SelfSupervisedLoss()(pc1, pc2, f1, vel)
    t2, items2 = R.SelfSupervisedLoss()(pc1, pc2, f2, vel)
    t1.backward(); t2.backward()
    assert set(items1) == set(items2)
    for k in items2:
        assert abs(items1[k] - items2[k]) <= 1e-4 * max(abs(items2[k]), 1e-6), (k, items1[k], items2[k])
    assert (f1.grad - f2.grad).abs().max() <= 1e-4 * f2.grad.abs().max()
