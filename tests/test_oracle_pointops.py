"""CPU checks of the C oracle's operator semantics (edge cases the reference's kernels define)."""
import torch

from oracle import pointops as P


def test_ball_query_semantics():
    xyz = torch.tensor([[[0., 0, 0], [1, 0, 0], [5, 0, 0], [0.5, 0, 0]]])
    q = torch.tensor([[[0., 0, 0], [100, 0, 0]]])
    idx = P.ball_query(1.5, 3, xyz, q)
    assert idx[0, 0].tolist() == [0, 1, 3]            # first 3 in index order
    assert idx[0, 1].tolist() == [0, 0, 0]            # no hit: row stays at the caller's zeros
    idx = P.ball_query(0.75, 4, xyz, q)
    assert idx[0, 0].tolist() == [0, 3, 0, 0]         # padded with the FIRST hit
    idx = P.ball_query(1.0, 2, xyz, q)                # strict <: the point at distance exactly 1 is out
    assert idx[0, 0].tolist() == [0, 3]


def test_knn_orders_ties_by_index_and_handles_short_sets():
    known = torch.tensor([[[1., 0, 0], [0, 1, 0], [0, 0, 1], [2, 0, 0]]])
    unk = torch.zeros(1, 1, 3)
    d2, idx = P.knn(3, unk, known)
    assert idx[0, 0].tolist() == [0, 1, 2] and d2[0, 0].tolist() == [1, 1, 1]
    d2, idx = P.knn(6, unk, known)                    # fewer candidates than k: (1e40 -> inf, 0)
    assert idx[0, 0].tolist() == [0, 1, 2, 3, 0, 0] and torch.isinf(d2[0, 0, 4:]).all()


def test_knn_point_matches_torch_topk_sets():
    g = torch.Generator().manual_seed(0)
    xyz = torch.rand(3, 200, 3, generator=g) * 40
    q = torch.rand(3, 150, 3, generator=g) * 40
    idx, d = P.knn_point(8, xyz, q)
    dist = -2 * q @ xyz.transpose(1, 2) + (q ** 2).sum(-1)[..., None] + (xyz ** 2).sum(-1)[:, None]
    ref = dist.clamp_min(0).topk(8, largest=False)[1].sort(-1)[0]
    assert torch.equal(idx.long().sort(-1)[0], ref)
    assert (d[..., 1:] >= d[..., :-1]).all()


def test_fps_matches_naive_argmax_when_untied():
    g = torch.Generator().manual_seed(1)
    xyz = torch.rand(2, 300, 3, generator=g)
    idx = P.furthest_point_sample(xyz, 16)
    for b in range(2):
        sel = [0]
        dist = torch.full((300,), 1e10)
        for _ in range(15):
            dist = torch.minimum(dist, ((xyz[b] - xyz[b, sel[-1]]) ** 2).sum(-1))
            sel.append(int(dist.argmax()))
        assert idx[b].tolist() == sel


def test_group_gather_interpolate_roundtrip():
    g = torch.Generator().manual_seed(2)
    pts = torch.rand(2, 5, 30, generator=g)
    idx = torch.randint(0, 30, (2, 7, 4), generator=g, dtype=torch.int32)
    out = P.group_points(pts, idx)
    assert torch.equal(out, torch.gather(pts, 2, idx.long().view(2, 1, -1).expand(2, 5, -1)).view(2, 5, 7, 4))
    grad = P.group_points_grad(torch.ones_like(out), idx, 30)
    cnt = torch.zeros(2, 30)
    for b in range(2):
        cnt[b] = torch.bincount(idx[b].flatten().long(), minlength=30).float()
    assert torch.equal(grad, cnt[:, None, :].expand(2, 5, 30))
