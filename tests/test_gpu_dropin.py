"""GPU tests of the drop-in surface beyond the fixed-size batched forward: un-resampled clouds (N1 != N2, one pair per call, a different
size every call -- the reference's evaluation loop, main.py:203 / main_util.py:118-145 / dataset/vod.py:92-93), duplicate-padded clouds
(vod.py:102-110), mode='train' with pseudo labels (models/cmflow.py:181-182), ill-conditioned Kabsch systems, the pipelined host entry
points, CUDA-graph replay, and engines that live on a device other than the current one."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from cmflow_b200._lib import CmfError, check, dptr, lib, stream_ptr   # noqa: E402
from cmflow_b200.cmflow import CMFlow, CMFlow_T, RaFlow   # noqa: E402
from cmflow_b200.synth import make_padded_pairs, make_pairs, synthetic_state_dict   # noqa: E402
from oracle import cmflow_oracle as O   # noqa: E402
from tests.helpers import (case_inputs, case_weights, check_outputs, check_raflow_outputs, knn_sets_equal, load_golden,   # noqa: E402
                           rel_err)

DEV = "cuda"
PRECISIONS = ["fp32", "fp16x3"]


class Args:
    num_points = 256
    stat_thres = 0.5
    rigid_thres = 0.15


def cmflow(sd, precision, cls=CMFlow, dev=DEV):
    net = cls(Args()); net.load_state_dict(sd, strict=True); net = net.to(dev); net.set_precision(precision)
    return net


def run(net, inp, label_m=None, mode="test", g=None):
    pc1, pc2, ft1, ft2 = (t.to(next(net.parameters()).device) for t in inp[:4])
    with torch.no_grad():
        if isinstance(net, CMFlow_T):
            sf, cls, T, mask, g = net(pc1, pc2, ft1, ft2, label_m, mode, g)
        else:
            sf, cls, T, mask = net(pc1, pc2, ft1, ft2, label_m, mode)
    return {"sf_agg": sf.cpu(), "stat_cls": cls.cpu(), "pre_trans": T.cpu(), "mask": mask.cpu(), "gfeat": g}


def uid_sets(idx, mapping):
    B, N, k = idx.shape
    return torch.gather(mapping.long(), 1, idx.long().flatten(1)).view(B, N, k).sort(-1)[0]


@pytest.mark.parametrize("precision", PRECISIONS)
def test_duplicate_padded_clouds_match_reference_golden(golden_dir, precision):
    gold = load_golden(golden_dir, "cmflow_synth_padded_b2_n256.pt")
    meta = gold["meta"]
    net = cmflow(case_weights(meta, golden_dir), precision)
    inp, maps, _ = make_padded_pairs(meta["B"], meta["N"], meta["n_unique"], meta["data_seed"])
    out = run(net, inp)
    B, N = meta["B"], meta["N"]
    # exact distance ties between the copies of a point: the POINTS chosen must be the reference's, whichever copy stands for them
    assert torch.equal(uid_sets(net.tap("knn12", (B, N, 8), torch.int32).cpu(), maps[1]), gold["knn12_uid"].long())
    assert torch.equal(uid_sets(net.tap("knn11", (B, N, 8), torch.int32).cpu(), maps[0]), gold["knn11_uid"].long())
    assert rel_err(net.tap("prop", (B, N, 256)).cpu()[0, ::4].t(), gold["prop_sub"], per_pair=False) <= 1e-4
    print(precision, check_outputs(out, gold))
    # copies of a point are the same point: identical neighbourhoods, hence the same flow (up to the rounding of a different tile position)
    m1 = maps[0]
    first = torch.stack([torch.stack([(m1[b] == u).nonzero()[0, 0] for u in m1[b]]) for b in range(B)])      # first copy of every point
    assert rel_err(out["sf_agg"], torch.gather(out["sf_agg"], 2, first[:, None, :].expand(B, 3, N))) <= 1e-6


@pytest.mark.parametrize("precision", PRECISIONS)
def test_train_mode_pseudo_labels_match_reference_golden(golden_dir, precision):
    gold = load_golden(golden_dir, "cmflow_synth_train_b2_n256.pt")
    net = cmflow(case_weights(gold["meta"], golden_dir), precision)
    inp = case_inputs(gold["meta"])
    out = run(net, inp, label_m=gold["label_m"].to(DEV), mode="train")
    assert torch.equal(out["mask"], gold["mask"])
    assert (out["stat_cls"] - gold["stat_cls"]).abs().max() <= 1e-4
    assert rel_err(out["pre_trans"][:, :3], gold["pre_trans"][:, :3]) <= 1e-4
    assert rel_err(out["sf_agg"], gold["sf_agg"]) <= 1e-4
    # mode='test' ignores the labels (cmflow.py:181: both conditions), and so does label_m=None in train mode
    ref = run(net, inp)
    for o in (run(net, inp, label_m=gold["label_m"].to(DEV), mode="test"), run(net, inp, label_m=None, mode="train")):
        assert torch.equal(o["sf_agg"], ref["sf_agg"]) and torch.equal(o["mask"], ref["mask"])


@pytest.mark.parametrize("precision", PRECISIONS)
def test_real_radar_frames_n1_ne_n2_match_reference_golden(golden_dir, precision):
    """The evaluation loop's call pattern: B=1, real clouds, N1 != N2, a different size every call, ONE engine throughout."""
    gold = load_golden(golden_dir, "real_radar_ckpt_n1n2.pt")
    sd = case_weights(gold["meta"], golden_dir)
    sdr = case_weights({"weights": gold["meta"]["weights_raflow"], "model": "raflow"}, golden_dir)
    if sd is None or sdr is None:
        pytest.skip("reference checkpoints not available")
    net, netr = cmflow(sd, precision), cmflow(sdr, precision, RaFlow)
    frames = gold["frames"]
    for fr in list(frames) + list(reversed(frames)):               # sizes go up and down: workspace is reused, results must not depend on history
        n1, n2 = fr["pc1"].shape[2], fr["pc2"].shape[2]
        out = run(net, (fr["pc1"], fr["pc2"], fr["ft1"], fr["ft2"]))
        g = fr["cmflow"]
        assert knn_sets_equal(net.tap("knn12", (1, n1, 8), torch.int32).cpu(), g["knn12"])
        assert knn_sets_equal(net.tap("knn11", (1, n1, 8), torch.int32).cpu(), g["knn11"])
        assert rel_err(net.tap("prop", (1, n1, 256)).cpu()[0, ::4].t(), g["prop_sub"], per_pair=False) <= 1e-4
        print(precision, fr["source"], n1, n2, check_outputs(out, g))
        if fr["raflow"] is not None:
            with torch.no_grad():
                o, sf, T, ms = netr(*(fr[k].to(DEV) for k in ("pc1", "pc2", "ft1", "ft2")), fr["raflow"]["interval"].to(DEV))
            print(precision, "raflow", check_raflow_outputs({"output": o.cpu(), "sf_agg": sf.cpu(), "pre_trans": T.cpu(), "mask_s": ms.cpu()}, fr["raflow"]))


@pytest.mark.parametrize("precision", PRECISIONS)
def test_temporal_n1_ne_n2_matches_oracle(precision):
    sd = synthetic_state_dict(3, temporal=True)
    net = cmflow(sd, precision, CMFlow_T)
    g, gref = None, None
    for step, (n1, n2) in enumerate(((200, 256), (256, 173), (131, 131))):
        a, b = make_pairs(2, n1, seed=70 + step), make_pairs(2, n2, seed=80 + step)
        inp = (a[0], b[1], a[2], b[3])
        out = run(net, inp, g=g)
        ref = O.cmflow_forward(sd, *inp, temporal=True, gfeat_prev=gref)
        print(precision, n1, n2, check_outputs(out, ref))
        assert rel_err(out["gfeat"].cpu(), ref["gfeat"]) <= 1e-4
        g, gref = out["gfeat"], ref["gfeat"]


def test_host_entry_points_with_n1_ne_n2_and_pipelining():
    """forward_host with N1 != N2, and the two-slot submit / wait form: every call equals the device entry point on that call's inputs."""
    sd = synthetic_state_dict(0)
    net = cmflow(sd, "fp16x3")
    calls = []
    for seed, B, n1, n2 in ((1, 3, 256, 256), (2, 3, 256, 200), (3, 3, 256, 200), (4, 2, 180, 256), (5, 3, 256, 256), (6, 3, 256, 256)):
        a, b = make_pairs(B, n1, seed=seed), make_pairs(B, n2, seed=seed + 50)
        calls.append(tuple(t.pin_memory() for t in (a[0], b[1], a[2], b[3])))
    want = [run(net, c) for c in calls]
    for c, w in zip(calls, want):
        h = net.forward_host(*c)
        assert torch.equal(h["sf_agg"], w["sf_agg"]) and torch.equal(h["pre_trans"], w["pre_trans"]) and torch.equal(h["mask"].bool(), w["mask"])
    # pipelined: keep two calls in flight
    outs = [None] * len(calls)
    for i, c in enumerate(calls):
        slot = i & 1
        if i >= 2:
            net.wait_host(slot)
        outs[i] = net.submit_host(slot, *c)
    net.wait_host(0); net.wait_host(1)
    for h, w in zip(outs, want):
        assert torch.equal(h["sf_agg"], w["sf_agg"]) and torch.equal(h["stat_cls"], w["stat_cls"])
        assert torch.equal(h["pre_trans"], w["pre_trans"]) and torch.equal(h["mask"].bool(), w["mask"])
    with pytest.raises(CmfError):                                   # a slot in flight cannot be submitted to again
        net.submit_host(0, *calls[0]); net.submit_host(0, *calls[0])
    net.wait_host(0)


@pytest.mark.parametrize("precision,temporal", [("fp32", False), ("fp16x3", False), ("fp16x3", True)])
def test_host_graph_replay(monkeypatch, precision, temporal):
    """CMF_HOST_GRAPH=1: first call of a shape eager, second captured, later ones replayed; each equals the device entry point."""
    monkeypatch.setenv("CMF_HOST_GRAPH", "1")
    sd = synthetic_state_dict(3 if temporal else 0, temporal=temporal)
    net = cmflow(sd, precision, CMFlow_T if temporal else CMFlow)
    side = torch.cuda.Stream()
    g_dev, g_host = None, None
    for seed, B, N in ((1, 1, 256), (2, 1, 256), (3, 1, 256), (4, 2, 200), (5, 1, 256), (6, 2, 200), (7, 2, 200)):
        inp = make_pairs(B, N, seed=seed)
        if temporal and (g_dev is None or g_dev.shape[0] != B):
            g_dev, g_host = None, None
        dev = run(net, inp, g=g_dev)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            host = net.forward_host(*[t.pin_memory() for t in inp[:4]], gfeat=g_host)
        assert torch.equal(host["sf_agg"], dev["sf_agg"]) and torch.equal(host["pre_trans"], dev["pre_trans"]), (seed, B, N)
        assert torch.equal(host["stat_cls"], dev["stat_cls"]) and torch.equal(host["mask"].bool(), dev["mask"])
        if temporal:
            assert torch.equal(host["gfeat"], dev["gfeat"].cpu())
            g_dev, g_host = dev["gfeat"], host["gfeat"].clone()
    assert lib().cmf_model_host_graphs(net._handle) >= 1


def test_illconditioned_kabsch_matches_reference_golden(golden_dir):
    gold = load_golden(golden_dir, "kabsch_illcond_n128.pt")
    A, Bp, W = gold["A"].to(DEV), gold["B"].to(DEV), gold["W"].to(DEV)
    nb = A.shape[0]
    T = torch.empty(nb, 4, 4, device=DEV)
    check(lib().cmf_weighted_kabsch(nb, 128, dptr(A), dptr(Bp), dptr(W), dptr(T), stream_ptr()))
    e = rel_err(T.cpu()[:, :3], gold["T"][:, :3])
    print("ill-conditioned Kabsch vs torch.svd (fp32 reference):", e)
    assert e <= 1e-4


def test_rank_deficient_kabsch_is_optimal_and_orthonormal():
    """H of rank 2 and rank 1 (all weight on three / two points; exactly planar clouds): V U^T is not unique there and the reference's own
    answer depends on LAPACK's sign choices, so the check is on what IS defined -- R orthonormal, and the weighted residual no worse than
    that of torch.linalg.svd's solution in fp64."""
    g = torch.Generator().manual_seed(5)
    nb, N = 4, 128
    A = torch.randn(nb, 3, N, generator=g) * torch.tensor([20.0, 10.0, 1.0]).view(1, 3, 1)
    A[1, 2] = 0.0                                                 # exactly planar
    yaw = torch.tensor([0.2, -0.1, 0.05, 0.0])
    R = torch.zeros(nb, 3, 3)
    R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1], R[:, 2, 2] = yaw.cos(), -yaw.sin(), yaw.sin(), yaw.cos(), 1.0
    Bp = R @ A + torch.tensor([1.0, -0.2, 0.05]).view(1, 3, 1)
    W = torch.rand(nb, N, generator=g)
    W[0] = 0.0; W[0, 5:8] = 1.0                                    # three points: rank 2
    W[2] = 0.0; W[2, 40:42] = 1.0                                  # two points: rank 1
    W = W / W.sum(1, keepdim=True)
    T = torch.empty(nb, 4, 4, device=DEV)
    Ad, Bd, Wd = A.to(DEV), Bp.to(DEV), W.to(DEV)                  # keep the device copies alive across the asynchronous launch
    check(lib().cmf_weighted_kabsch(nb, N, dptr(Ad), dptr(Bd), dptr(Wd), dptr(T), stream_ptr()))
    T = T.cpu().double()
    Rg, tg = T[:, :3, :3], T[:, :3, 3:]
    assert torch.isfinite(T).all()
    assert (Rg @ Rg.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-6
    Tref, _ = O.weighted_kabsch(A.double(), Bp.double(), W.double())
    res = lambda Rm, tm: (W.double() * ((Rm @ A.double() + tm - Bp.double()) ** 2).sum(1)).sum(1)
    ours, ref = res(Rg, tg), res(Tref[:, :3, :3], Tref[:, :3, 3:])
    print("weighted residuals ours / torch.svd:", ours.tolist(), ref.tolist())
    assert (ours <= ref + 1e-6).all()


def test_input_validation_raises():
    net = cmflow(synthetic_state_dict(0), "fp32")
    pc1, pc2, ft1, ft2, _ = (t.to(DEV) for t in make_pairs(2, 64, seed=1))
    with pytest.raises(CmfError):
        net(pc1, pc2[:1], ft1, ft2, None, "test")                      # batch mismatch
    with pytest.raises(CmfError):
        net(pc1, pc2, ft1[:, :, :32], ft2, None, "test")               # feature1 not like pc1
    with pytest.raises(CmfError):
        net(pc1.cpu(), pc2, ft1, ft2, None, "test")                    # no CPU path
    with pytest.raises(CmfError):
        net(pc1, pc2, ft1, ft2, torch.zeros(2, 63, device=DEV), "train")   # label_m of the wrong size
    with pytest.raises(CmfError):
        net.forward_host(pc1.cpu(), pc2.cpu(), ft1.cpu(), ft2.cpu(), out={"sf_agg": torch.empty(1)})
    nett = cmflow(synthetic_state_dict(3, temporal=True), "fp32", CMFlow_T)
    with pytest.raises(CmfError):
        nett(pc1, pc2, ft1, ft2, None, "test", torch.zeros(2, 128, device=DEV))   # gfeat not (B,256)
    with pytest.warns(RuntimeWarning):
        net.train(); net(pc1, pc2, ft1, ft2, None, "test")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_engine_on_non_current_device():
    """The model lives on cuda:1 while cuda:0 is current (single-process multi-GPU, the reference's nn.DataParallel habit): weight
    tiling, workspace and launches must land on cuda:1."""
    sd = synthetic_state_dict(0)
    inp = make_pairs(2, 256, seed=3)
    torch.cuda.set_device(0)
    want = run(cmflow(sd, "fp16x3", dev="cuda:0"), inp)
    net1 = CMFlow(Args()); net1.load_state_dict(sd); net1 = net1.to("cuda:1")
    assert torch.cuda.current_device() == 0
    net1.set_precision("fp16x3")                                       # before the engine exists
    got = run(net1, inp)
    net1.set_precision("fp32"); net1.set_precision("fp16x3")           # and on a live engine, cuda:0 still current
    got2 = run(net1, inp)
    assert torch.cuda.current_device() == 0
    for k in ("sf_agg", "stat_cls", "pre_trans", "mask"):
        assert torch.equal(got[k], want[k]) and torch.equal(got2[k], want[k]), k
