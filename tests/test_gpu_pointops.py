"""GPU parity of Part 1 of the C ABI (the pointnet2_cuda operator set) -- bit-exact against
 (a) the C oracle and (b) the reference's own lib/src CUDA kernels compiled unmodified (oracle/_ref)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from cmflow_b200 import pointnet2_utils as PU   # noqa: E402
from cmflow_b200._lib import CmfError, dptr, lib, stream_ptr, check   # noqa: E402
from oracle import pointops as P   # noqa: E402
from oracle import refcuda as R    # noqa: E402

DEV = "cuda"


def cloud(B, N, seed, scale=(50.0, 20.0, 2.0), dup=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, N, 3, generator=g) * torch.tensor(scale) * torch.tensor([1.0, 2.0, 2.0]) - torch.tensor([0.0, scale[1], scale[2]])
    if dup:                                            # padding-by-resampling duplicates (dataset/vod.py:102-104)
        src = torch.randint(0, N - dup, (B, dup), generator=g)
        x[:, N - dup:] = torch.gather(x, 1, src[..., None].expand(B, dup, 3))
    return x.contiguous()


SHAPES = [(2, 256, 256), (3, 200, 77), (1, 33, 40), (2, 4096, 300), (1, 1, 5), (2, 3000, 3000)]


@pytest.mark.parametrize("B,N,M", SHAPES)
@pytest.mark.parametrize("radius,nsample", [(2.0, 4), (4.0, 8), (8.0, 16), (16.0, 32), (0.5, 5), (1000.0, 64)])
def test_ball_query_bit_exact(B, N, M, radius, nsample):
    xyz, new = cloud(B, N, 1, dup=min(8, N // 4)), cloud(B, M, 2)
    if M == N:
        new = xyz.clone()
    want = P.ball_query(radius, nsample, xyz, new)
    got = PU.ball_query(radius, nsample, xyz.to(DEV), new.to(DEV))
    assert torch.equal(got.cpu(), want)
    if R.available():
        assert torch.equal(got, R.ball_query(radius, nsample, xyz.to(DEV), new.to(DEV)))


@pytest.mark.parametrize("B,N,M", SHAPES)
@pytest.mark.parametrize("k", [1, 3, 8, 16, 32, 40])
def test_knn_bit_exact(B, N, M, k):
    known, unk = cloud(B, N, 3, dup=min(8, N // 4)), cloud(B, M, 4)
    d2w, iw = P.knn(k, unk, known)
    from cmflow_b200 import pointnet2_cuda as K
    d2 = torch.empty(B, M, k, device=DEV)
    idx = torch.empty(B, M, k, dtype=torch.int32, device=DEV)
    K.knn_wrapper(B, M, N, k, unk.to(DEV), known.to(DEV), d2, idx)          # raw kernel output: squared distances
    assert torch.equal(idx.cpu(), iw)
    assert torch.equal(d2.cpu(), d2w)
    dist, idx2 = PU.knn(k, unk.to(DEV), known.to(DEV))                      # KNN.forward returns sqrt (pointnet2_utils.py:97)
    assert torch.equal(idx2, idx) and torch.equal(dist, torch.sqrt(d2))
    if R.available():
        d2r, ir = R.knn(k, unk.to(DEV), known.to(DEV))
        assert torch.equal(idx, ir) and torch.equal(d2, d2r)


@pytest.mark.parametrize("B,N,M", SHAPES)
def test_three_nn_and_interpolate_bit_exact(B, N, M):
    known, unk = cloud(B, N, 5), cloud(B, M, 6)
    d2w, iw = P.three_nn(unk, known)
    dist, idx = PU.three_nn(unk.to(DEV), known.to(DEV))
    assert torch.equal(idx.cpu(), iw) and torch.equal(dist, torch.sqrt(d2w.to(DEV)))
    g = torch.Generator().manual_seed(7)
    feats = torch.randn(B, 19, N, generator=g)
    w = torch.rand(B, M, 3, generator=g)
    want = P.three_interpolate(feats, iw, w)
    got = PU.three_interpolate(feats.to(DEV), idx, w.to(DEV))
    assert torch.equal(got.cpu(), want)
    if R.available():
        assert torch.equal(got, R.three_interpolate(feats.to(DEV), idx, w.to(DEV)))
        d2r, ir = R.three_nn(unk.to(DEV), known.to(DEV))
        assert torch.equal(idx, ir) and torch.equal(d2r.cpu(), d2w)


@pytest.mark.parametrize("B,C,N,Pn,S", [(2, 6, 256, 256, 4), (3, 1027, 200, 200, 8), (1, 64, 4096, 4096, 32), (2, 3, 40, 7, 5)])
def test_group_and_gather_exact(B, C, N, Pn, S):
    g = torch.Generator().manual_seed(8)
    pts = torch.randn(B, C, N, generator=g)
    idx = torch.randint(0, N, (B, Pn, S), generator=g, dtype=torch.int32)
    got = PU.grouping_operation(pts.to(DEV), idx.to(DEV))
    assert torch.equal(got.cpu(), P.group_points(pts, idx))
    if R.available():
        assert torch.equal(got, R.group_points(pts.to(DEV), idx.to(DEV)))
    idx1 = idx[:, :, 0].contiguous()
    got1 = PU.gather_operation(pts.to(DEV), idx1.to(DEV))
    assert torch.equal(got1.cpu(), P.gather_points(pts, idx1))
    if R.available():
        assert torch.equal(got1, R.gather_points(pts.to(DEV), idx1.to(DEV)))


def test_backward_kernels_match_oracle():
    g = torch.Generator().manual_seed(9)
    B, C, N, Pn, S = 2, 16, 100, 60, 8
    pts = torch.randn(B, C, N, generator=g).to(DEV).requires_grad_()
    idx = torch.randint(0, N, (B, Pn, S), generator=g, dtype=torch.int32)
    go = torch.randn(B, C, Pn, S, generator=g)
    PU.grouping_operation(pts, idx.to(DEV)).backward(go.to(DEV))
    torch.testing.assert_close(pts.grad.cpu(), P.group_points_grad(go, idx, N), rtol=1e-5, atol=1e-5)   # atomics: order differs
    pts.grad = None
    idx1 = idx[:, :, 0].contiguous()
    go1 = torch.randn(B, C, Pn, generator=g)
    PU.gather_operation(pts, idx1.to(DEV)).backward(go1.to(DEV))
    torch.testing.assert_close(pts.grad.cpu(), P.gather_points_grad(go1, idx1, N), rtol=1e-5, atol=1e-5)
    pts.grad = None
    idx3 = torch.randint(0, N, (B, Pn, 3), generator=g, dtype=torch.int32)
    w = torch.rand(B, Pn, 3, generator=g)
    PU.three_interpolate(pts, idx3.to(DEV), w.to(DEV)).backward(go1.to(DEV))
    torch.testing.assert_close(pts.grad.cpu(), P.three_interpolate_grad(go1, idx3, w, N), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,N,m", [(2, 256, 64), (1, 300, 128), (2, 1000, 33), (1, 20, 20), (1, 2500, 50), (2, 5, 3)])
def test_fps_bit_exact(B, N, m):
    xyz = cloud(B, N, 10, dup=min(6, N // 3))          # duplicates force ties through the reference's tree rule
    want = P.furthest_point_sample(xyz, m)
    got = PU.furthest_point_sample(xyz.to(DEV), m)
    assert torch.equal(got.cpu(), want)
    if R.available():
        ref_idx, _ = R.furthest_point_sample(xyz.to(DEV), m)
        assert torch.equal(got, ref_idx)


@pytest.mark.parametrize("B,N,S", [(2, 256, 256), (3, 200, 77), (2, 4096, 512), (1, 8, 8)])
@pytest.mark.parametrize("k", [8, 1, 32])
def test_knn_point_bit_exact(B, N, S, k):
    if k > N:
        pytest.skip("k > N")
    xyz, q = cloud(B, N, 11, dup=min(8, N // 4)), cloud(B, S, 12)
    iw, dw = P.knn_point(k, xyz, q)
    idx = torch.empty(B, S, k, dtype=torch.int32, device=DEV)
    dist = torch.empty(B, S, k, device=DEV)
    xd, qd = xyz.to(DEV), q.to(DEV)                    # keep the device tensors alive across the raw-pointer call
    check(lib().cmf_knn_point(B, N, S, k, dptr(xd), dptr(qd), dptr(idx), dptr(dist), stream_ptr()))
    assert torch.equal(idx.cpu(), iw) and torch.equal(dist.cpu(), dw)


def test_ball_query_ms_matches_four_single_queries():
    for B, N in ((2, 256), (1, 77), (2, 4096)):
        xyz = cloud(B, N, 13, dup=min(8, N // 4))
        planar = xyz.permute(0, 2, 1).contiguous().to(DEV)
        idx60 = torch.zeros(B, N, 60, dtype=torch.int32, device=DEV)
        check(lib().cmf_ball_query_ms(B, N, dptr(planar), dptr(idx60), stream_ptr()))
        want = torch.cat([P.ball_query(r, k, xyz, xyz) for r, k in ((2.0, 4), (4.0, 8), (8.0, 16), (16.0, 32))], -1)
        assert torch.equal(idx60.cpu(), want)


def test_errors_are_loud_not_fatal():
    x = torch.zeros(1, 4, 3)
    with pytest.raises(CmfError):
        PU.ball_query(1.0, 2, x, x)                     # CPU tensors: no CPU path
    with pytest.raises(CmfError):
        PU.knn(300, x.to(DEV), x.to(DEV))               # k > 200
    xt = torch.zeros(1, 3, 8, device=DEV).transpose(1, 2)
    with pytest.raises((CmfError, AssertionError)):
        PU.ball_query(1.0, 2, xt, xt)                   # non-contiguous


def test_query_and_group_matches_reference_composition():
    xyz = cloud(2, 256, 14)
    g = torch.Generator().manual_seed(15)
    feats = torch.randn(2, 5, 256, generator=g)
    out = PU.QueryAndGroup(4.0, 8)(xyz.to(DEV), xyz.to(DEV), feats.to(DEV))
    idx = P.ball_query(4.0, 8, xyz, xyz)
    gx = P.group_points(xyz.transpose(1, 2).contiguous(), idx) - xyz.transpose(1, 2).unsqueeze(-1)
    want = torch.cat([gx, P.group_points(feats, idx)], 1)
    assert torch.equal(out.cpu(), want)


@pytest.mark.parametrize("case", ["all_equal", "grid_ties", "few_candidates", "inf_nan", "chunk_boundary"])
def test_knn_adversarial_order_and_ties(case):
    """The warp-distributed k-NN list against the C restatement (and the reference's own kernel) where ORDER is decided by ties:
    identical points (every distance equal: index order), integer-grid clouds (many exact ties at every rank), fewer candidates than k
    (unfilled slots -> index 0), non-finite coordinates (never selected), and candidates that straddle the 2048-point staging chunk."""
    g = torch.Generator().manual_seed(5)
    B = 2
    if case == "all_equal":
        known = torch.ones(B, 300, 3) * torch.tensor([3.0, -2.0, 0.5]); unk = torch.randn(B, 70, 3, generator=g)
    elif case == "grid_ties":
        known = torch.randint(-3, 4, (B, 500, 3), generator=g).float(); unk = torch.randint(-3, 4, (B, 130, 3), generator=g).float()
    elif case == "few_candidates":
        known = torch.randn(B, 5, 3, generator=g); unk = torch.randn(B, 33, 3, generator=g)
    elif case == "inf_nan":
        known = torch.randn(B, 200, 3, generator=g)
        known[0, 3, 0] = float("inf"); known[0, 40, 1] = float("nan"); known[1, 0, 2] = float("-inf"); known[1, 199] = float("nan")
        unk = torch.randn(B, 64, 3, generator=g)
    else:
        known = torch.randn(B, 2048 + 37, 3, generator=g) * 5
        known[:, 2040:2060] = known[:, :20]                      # exact duplicates on both sides of the chunk boundary
        unk = known[:, :40].clone()
    from cmflow_b200 import pointnet2_cuda as K
    N, M = known.shape[1], unk.shape[1]
    for k in (1, 3, 8, 32):
        d2w, iw = P.knn(k, unk, known)
        d2 = torch.empty(B, M, k, device=DEV); idx = torch.empty(B, M, k, dtype=torch.int32, device=DEV)
        K.knn_wrapper(B, M, N, k, unk.to(DEV), known.to(DEV), d2, idx)
        assert torch.equal(idx.cpu(), iw), (case, k)
        assert torch.equal(d2.cpu().nan_to_num(posinf=1e30), d2w.nan_to_num(posinf=1e30)), (case, k)
        if R.available():
            d2r, ir = R.knn(k, unk.to(DEV), known.to(DEV))
            assert torch.equal(idx, ir), (case, k, "reference kernel")
        if k <= N and case != "inf_nan":                         # the model's k-NN (expanded-form distance, radarflow_util.py:88-99)
            ipw, dpw = P.knn_point(k, known, unk)
            ip = torch.empty(B, M, k, dtype=torch.int32, device=DEV); dp = torch.empty(B, M, k, device=DEV)
            kd, ud = known.to(DEV), unk.to(DEV)
            check(lib().cmf_knn_point(B, N, M, k, dptr(kd), dptr(ud), dptr(ip), dptr(dp), stream_ptr()))
            assert torch.equal(ip.cpu(), ipw) and torch.equal(dp.cpu(), dpw), (case, k, "knn_point")
