"""Host logic of cmflow_b200/pointnet2_utils.py (shapes, caller-allocated outputs, autograd wiring) on CPU: the module's launcher table is
swapped for the oracle's CPU restatement of the same ten wrappers (oracle/pointops.py:as_pointnet2_module), so every operator and every
backward runs end to end without a GPU and is compared with direct oracle calls / torch autograd."""
import pytest
import torch

from oracle import pointops as P


@pytest.fixture()
def PU(monkeypatch):
    from cmflow_b200 import pointnet2_utils as mod
    monkeypatch.setattr(mod, "_k", P.as_pointnet2_module())
    return mod


def _cloud(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, N, 3, generator=g) * torch.tensor([50.0, 40.0, 4.0])).contiguous()


def test_index_operators(PU):
    xyz, q = _cloud(2, 96, 1), _cloud(2, 40, 2)
    assert torch.equal(PU.ball_query(6.0, 8, xyz, q), P.ball_query(6.0, 8, xyz, q))
    assert torch.equal(PU.furthest_point_sample(xyz, 17), P.furthest_point_sample(xyz, 17))
    d2, idx = P.knn(5, q, xyz)
    dist, got = PU.knn(5, q, xyz)
    assert torch.equal(got, idx) and torch.equal(dist, d2.sqrt()) and got.dtype == torch.int32
    d3, i3 = P.three_nn(q, xyz)
    dist3, got3 = PU.three_nn(q, xyz)
    assert torch.equal(got3, i3) and torch.equal(dist3, d3.sqrt())
    # index outputs are not differentiable and their backward returns one None per forward input
    xr = xyz.clone().requires_grad_(True)
    dist, got = PU.knn(5, q, xr)
    assert not got.requires_grad


def test_gather_group_interpolate_and_their_gradients(PU):
    g = torch.Generator().manual_seed(3)
    B, C, N = 2, 6, 50
    feats = torch.randn(B, C, N, generator=g).contiguous()
    idx = torch.randint(0, N, (B, 11), generator=g).int()
    f = feats.clone().requires_grad_(True)
    out = PU.gather_operation(f, idx)
    want = torch.gather(feats, 2, idx.long().unsqueeze(1).expand(B, C, 11))
    assert torch.equal(out, want)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    fr = feats.clone().requires_grad_(True)
    (torch.gather(fr, 2, idx.long().unsqueeze(1).expand(B, C, 11)) * w).sum().backward()
    assert torch.allclose(f.grad, fr.grad, atol=1e-6)

    gidx = torch.randint(0, N, (B, 9, 4), generator=g).int()
    f = feats.clone().requires_grad_(True)
    grp = PU.grouping_operation(f, gidx)
    ix = gidx.long().view(B, 1, 36).expand(B, C, 36)
    assert torch.equal(grp, torch.gather(feats, 2, ix).view(B, C, 9, 4))
    w = torch.randn_like(grp)
    (grp * w).sum().backward()
    fr = feats.clone().requires_grad_(True)
    (torch.gather(fr, 2, ix).view(B, C, 9, 4) * w).sum().backward()
    assert torch.allclose(f.grad, fr.grad, atol=1e-6)

    i3 = torch.randint(0, N, (B, 13, 3), generator=g).int()
    w3 = torch.rand(B, 13, 3, generator=g)
    w3 = (w3 / w3.sum(-1, keepdim=True)).contiguous()
    f = feats.clone().requires_grad_(True)
    mixed = PU.three_interpolate(f, i3, w3)
    ref = (torch.gather(feats, 2, i3.long().view(B, 1, 39).expand(B, C, 39)).view(B, C, 13, 3) * w3.unsqueeze(1)).sum(-1)
    assert torch.allclose(mixed, ref, atol=1e-6)
    wo = torch.randn_like(mixed)
    (mixed * wo).sum().backward()
    fr = feats.clone().requires_grad_(True)
    ((torch.gather(fr, 2, i3.long().view(B, 1, 39).expand(B, C, 39)).view(B, C, 13, 3) * w3.unsqueeze(1)).sum(-1) * wo).sum().backward()
    assert torch.allclose(f.grad, fr.grad, atol=1e-5)


def test_query_and_group_and_group_all(PU):
    xyz, q = _cloud(2, 64, 5), _cloud(2, 20, 6)
    feats = torch.randn(2, 4, 64, generator=torch.Generator().manual_seed(7)).contiguous()
    qg = PU.QueryAndGroup(8.0, 6)
    out = qg(xyz, q, feats)
    idx = P.ball_query(8.0, 6, xyz, q)
    ix = idx.long().view(2, 1, 120)
    rel = torch.gather(xyz.transpose(1, 2), 2, ix.expand(2, 3, 120)).view(2, 3, 20, 6) - q.transpose(1, 2).unsqueeze(-1)
    assert out.shape == (2, 7, 20, 6) and torch.equal(out[:, :3], rel)
    assert torch.equal(out[:, 3:], torch.gather(feats, 2, ix.expand(2, 4, 120)).view(2, 4, 20, 6))
    assert torch.equal(PU.QueryAndGroup(8.0, 6, use_xyz=False)(xyz, q, feats), out[:, 3:])
    assert torch.equal(qg(xyz, q), rel)
    ga = PU.GroupAll()
    assert ga(xyz, None, feats).shape == (2, 7, 1, 64) and torch.equal(ga(xyz, None), xyz.transpose(1, 2).unsqueeze(2))
