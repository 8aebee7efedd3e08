"""GPU parity of the whole-forward engine (Part 3 of the C ABI) against the golden vectors produced by the
unmodified reference, the CPU oracle, and the stage-by-stage emulator."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from cmflow_b200 import weights   # noqa: E402
from cmflow_b200._lib import check, dptr, lib, stream_ptr   # noqa: E402
from cmflow_b200.cmflow import CMFlow, CMFlow_T   # noqa: E402
from cmflow_b200.synth import make_pairs, synthetic_state_dict   # noqa: E402
from oracle import cmflow_oracle as O   # noqa: E402
from tests.helpers import case_inputs, case_weights, check_outputs, knn_sets_equal, load_golden, rel_err   # noqa: E402
from tests.pipeline_emulator import emulate   # noqa: E402

DEV = "cuda"


class Args:
    num_points = 256
    stat_thres = 0.5


def build(meta, golden_dir):
    sd = case_weights(meta, golden_dir)
    if sd is None:
        pytest.skip("reference checkpoint not available")
    net = (CMFlow_T if meta["model"] == "cmflow_t" else CMFlow)(Args())
    net.load_state_dict(sd, strict=True)
    return net.to(DEV), sd


def run(net, inp, g=None):
    pc1, pc2, ft1, ft2 = (t.to(DEV) for t in inp[:4])
    with torch.no_grad():
        if isinstance(net, CMFlow_T):
            sf, cls, T, mask, g = net(pc1, pc2, ft1, ft2, None, "test", g)
        else:
            sf, cls, T, mask = net(pc1, pc2, ft1, ft2, None, "test")
    return {"sf_agg": sf.cpu(), "stat_cls": cls.cpu(), "pre_trans": T.cpu(), "mask": mask.cpu(), "gfeat": g}


CASES = ["cmflow_synth_b2_n256.pt", "cmflow_synth_w1_b2_n256.pt", "cmflow_synth_b3_n200.pt", "cmflow_synth_b2_n40.pt",
         "cmflow_ckpt_b2_n256.pt"]


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference_golden(golden_dir, name):
    gold = load_golden(golden_dir, name)
    net, sd = build(gold["meta"], golden_dir)
    inp = case_inputs(gold["meta"])
    out = run(net, inp)
    B, N = gold["meta"]["B"], gold["meta"]["N"]
    # integer work: neighbour sets identical to the reference's own torch.topk sets
    assert knn_sets_equal(net.tap("knn12", (B, N, 8), torch.int32).cpu(), gold["knn12"])
    assert knn_sets_equal(net.tap("knn11", (B, N, 8), torch.int32).cpu(), gold["knn11"])
    E = net.tap("E", (B, N, 800)).cpu()
    assert rel_err(E[0, ::4, 0:256].t(), gold["f1_sub"], per_pair=False) <= 1e-4
    assert rel_err(E[0, ::4, 256:768].t(), gold["cor_sub"], per_pair=False) <= 1e-4
    assert rel_err(net.tap("f2", (B, N, 256)).cpu()[0, ::4].t(), gold["f2_sub"], per_pair=False) <= 1e-4
    assert rel_err(net.tap("prop", (B, N, 256)).cpu()[0, ::4].t(), gold["prop_sub"], per_pair=False) <= 1e-4
    errs = check_outputs(out, gold)
    print(name, errs)
    assert out["sf_agg"].shape == (B, 3, N) and out["stat_cls"].shape == (B, 1, N)
    assert out["pre_trans"].shape == (B, 4, 4) and out["mask"].dtype == torch.bool


def test_temporal_clip_matches_reference_golden(golden_dir):
    gold = load_golden(golden_dir, "cmflow_t_synth_b2_n256.pt")
    net, sd = build(gold["meta"], golden_dir)
    inp = case_inputs(gold["meta"])
    g = None
    for step in gold["steps"]:
        out = run(net, inp, g)
        check_outputs(out, step)
        assert rel_err(out["gfeat"].cpu(), step["gfeat"]) <= 1e-4
        g = out["gfeat"]


def test_stage_taps_match_fp64_emulation(golden_dir):
    """Every stage boundary vs an fp64 replay of the same pipeline on the same packed weights: isolates
    kernel arithmetic error (summation order only) from everything else."""
    gold = load_golden(golden_dir, "cmflow_synth_b2_n256.pt")
    net, sd = build(gold["meta"], golden_dir)
    inp = case_inputs(gold["meta"])
    out = run(net, inp)
    B, N = 2, 256
    em = emulate(weights.pack(sd, False), *inp[:4], dtype=torch.float64)
    assert torch.equal(net.tap("bq1", (B, N, 60), torch.int32).cpu().long(), em["bq1"])
    assert torch.equal(net.tap("bq2", (B, N, 60), torch.int32).cpu().long(), em["bq2"])
    assert torch.equal(net.tap("knn12", (B, N, 8), torch.int32).cpu().long(), em["knn12"])
    assert torch.equal(net.tap("knn11", (B, N, 8), torch.int32).cpu().long(), em["knn11"])
    E = net.tap("E", (B, N, 800)).cpu()
    for name, got, want in (("f1", E[..., 0:256], em["f1"]), ("f2", net.tap("f2", (B, N, 256)).cpu(), em["f2"]),
                            ("cor", E[..., 256:768], em["cor"]), ("prop", net.tap("prop", (B, N, 256)).cpu(), em["prop"]),
                            ("flow", net.tap("flow", (B, 3, N)).cpu(), em["flow"])):
        e = rel_err(got, want)
        print(name, e)
        assert e <= 2e-5, (name, e)
    assert (out["stat_cls"].double() - em["stat_cls"]).abs().max() <= 2e-5


def test_chunked_batch_equals_single_chunk(golden_dir, monkeypatch):
    gold = load_golden(golden_dir, "cmflow_synth_b3_n200.pt")
    net, sd = build(gold["meta"], golden_dir)
    inp = case_inputs(gold["meta"])
    whole = run(net, inp)
    monkeypatch.setenv("CMF_CHUNK_PAIRS", "1")
    net2, _ = build(gold["meta"], golden_dir)
    parts = run(net2, inp)
    for k in ("sf_agg", "stat_cls", "pre_trans", "mask"):
        assert torch.equal(whole[k], parts[k]), k            # pairs are independent: bitwise identical


def test_host_entry_point_equals_device_entry_point(golden_dir):
    gold = load_golden(golden_dir, "cmflow_synth_b2_n256.pt")
    net, sd = build(gold["meta"], golden_dir)
    inp = case_inputs(gold["meta"])
    dev = run(net, inp)
    host = net.forward_host(*[t.pin_memory() for t in inp[:4]])
    assert torch.equal(host["sf_agg"], dev["sf_agg"]) and torch.equal(host["stat_cls"], dev["stat_cls"])
    assert torch.equal(host["pre_trans"], dev["pre_trans"]) and torch.equal(host["mask"].bool(), dev["mask"])


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_host_entry_point_graph_replay(monkeypatch, precision):
    """With CMF_HOST_GRAPH=1 cmf_model_forward_host runs a shape eagerly once, captures its kernel sequence as a CUDA graph on the second call
    (on a capturable, i.e. non-legacy-default, stream) and replays it afterwards: every call must equal the device entry point on that call's
    inputs.  A shape change re-allocates the workspace and drops the cached graphs."""
    monkeypatch.setenv("CMF_HOST_GRAPH", "1")
    net = CMFlow(Args()); net.load_state_dict(synthetic_state_dict(0)); net = net.to(DEV); net.set_precision(precision)
    side = torch.cuda.Stream()
    for seed, B, N in ((1, 3, 256), (2, 3, 256), (3, 3, 256), (4, 2, 200), (5, 3, 256), (6, 2, 200), (7, 2, 200)):
        inp = make_pairs(B, N, seed=seed)
        dev = run(net, inp)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            host = net.forward_host(*[t.pin_memory() for t in inp[:4]])
        assert torch.equal(host["sf_agg"], dev["sf_agg"]) and torch.equal(host["pre_trans"], dev["pre_trans"]), (seed, B, N)
        assert torch.equal(host["stat_cls"], dev["stat_cls"]) and torch.equal(host["mask"].bool(), dev["mask"])
    assert lib().cmf_model_host_graphs(net._handle) >= 1


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_host_entry_point_graph_replay_temporal(monkeypatch, precision):
    """The same for CMFlow-T with its carried GRU state: the replayed graph must follow the state fed to each call."""
    monkeypatch.setenv("CMF_HOST_GRAPH", "1")
    nett = CMFlow_T(Args()); nett.load_state_dict(synthetic_state_dict(3, temporal=True)); nett = nett.to(DEV); nett.set_precision(precision)
    side = torch.cuda.Stream()
    inp = make_pairs(2, 256, seed=9)
    for rep in range(2):
        g_dev, g_host = None, None
        for step in range(3):
            dev = run(nett, inp, g_dev)
            torch.cuda.synchronize()
            with torch.cuda.stream(side):
                host = nett.forward_host(*[t.pin_memory() for t in inp[:4]], gfeat=g_host)
            assert torch.equal(host["sf_agg"], dev["sf_agg"]) and torch.equal(host["gfeat"], dev["gfeat"].cpu()), (rep, step)
            g_dev, g_host = dev["gfeat"], host["gfeat"].clone()
    assert lib().cmf_model_host_graphs(nett._handle) >= 1


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_host_entry_point_repeated_calls(precision):
    """The host entry point over repeated calls, changing shapes and a carried CMFlow-T state (default: eager launches)."""
    net = CMFlow(Args()); net.load_state_dict(synthetic_state_dict(0)); net = net.to(DEV); net.set_precision(precision)
    for seed, B, N in ((1, 3, 256), (2, 3, 256), (4, 2, 200), (5, 3, 256)):
        inp = make_pairs(B, N, seed=seed)
        dev = run(net, inp)
        host = net.forward_host(*[t.pin_memory() for t in inp[:4]])
        assert torch.equal(host["sf_agg"], dev["sf_agg"]) and torch.equal(host["pre_trans"], dev["pre_trans"]), (seed, B, N)
    nett = CMFlow_T(Args()); nett.load_state_dict(synthetic_state_dict(3, temporal=True)); nett = nett.to(DEV); nett.set_precision(precision)
    inp = make_pairs(2, 256, seed=9)
    for rep in range(2):
        g_dev, g_host = None, None
        for step in range(3):
            dev = run(nett, inp, g_dev)
            host = nett.forward_host(*[t.pin_memory() for t in inp[:4]], gfeat=g_host)
            assert torch.equal(host["sf_agg"], dev["sf_agg"]) and torch.equal(host["gfeat"], dev["gfeat"].cpu()), (rep, step)
            g_dev, g_host = dev["gfeat"], host["gfeat"].clone()


def test_kabsch_matches_reference_golden(golden_dir):
    gold = load_golden(golden_dir, "kabsch_n128.pt")
    A, Bp, W = gold["A"].to(DEV), gold["B"].to(DEV), gold["W"].to(DEV)
    T = torch.empty(4, 4, 4, device=DEV)
    check(lib().cmf_weighted_kabsch(4, 128, dptr(A), dptr(Bp), dptr(W), dptr(T), stream_ptr()))
    assert rel_err(T.cpu()[:, :3], gold["T"][:, :3]) <= 1e-4
    assert torch.equal(T.cpu()[:, 3], torch.tensor([0., 0, 0, 1]).expand(4, 4))
    assert torch.linalg.det(T[3, :3, :3].cpu()) > 0.99          # reflected case handled as the reference does


def test_batch256_properties():
    """BASELINE.json configs[1] size (N=256, batch=256): size-independent properties instead of an oracle run --
    rotations orthonormal with det +1, static points carry exactly the rigid flow, outputs finite, and the first
    two pairs equal a B=2 run of the same inputs (batch independence)."""
    sd = synthetic_state_dict(0)
    net = CMFlow(Args()); net.load_state_dict(sd); net = net.to(DEV)
    inp = make_pairs(256, 256, seed=7)
    out = run(net, inp)
    R = out["pre_trans"][:, :3, :3].double()
    assert (R @ R.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-5
    assert (torch.linalg.det(R) - 1).abs().max() < 1e-5
    assert all(torch.isfinite(v).all() for k, v in out.items() if k in ("sf_agg", "stat_cls", "pre_trans"))
    pc1 = inp[0].double()
    rigid = (out["pre_trans"].double()[:, :3, :3] @ pc1 + out["pre_trans"].double()[:, :3, 3:]) - pc1
    m = out["mask"].unsqueeze(1).expand(-1, 3, -1)
    assert (out["sf_agg"].double() - rigid)[m].abs().max() < 1e-3
    small = run(net, tuple(t[:2] for t in inp))
    for k in ("sf_agg", "stat_cls", "pre_trans", "mask"):
        assert torch.equal(small[k], out[k][:2]), k


def test_dense_cloud_config_runs_and_matches_oracle_prefix():
    """BASELINE.json configs[4] shape class (N=4096) at B=1: integer stages exact vs the C oracle, outputs finite."""
    from oracle import pointops as P
    sd = synthetic_state_dict(0)
    net = CMFlow(Args()); net.load_state_dict(sd); net = net.to(DEV)
    inp = make_pairs(1, 4096, seed=3, dense=True)
    out = run(net, inp)
    x1t = inp[0].permute(0, 2, 1).contiguous(); x2t = inp[1].permute(0, 2, 1).contiguous()
    want = torch.cat([P.ball_query(r, k, x1t, x1t) for r, k in ((2.0, 4), (4.0, 8), (8.0, 16), (16.0, 32))], -1)
    assert torch.equal(net.tap("bq1", (1, 4096, 60), torch.int32).cpu(), want)
    assert torch.equal(net.tap("knn12", (1, 4096, 8), torch.int32).cpu(), P.knn_point(8, x2t, x1t)[0])
    assert torch.isfinite(out["sf_agg"]).all()
    # the whole forward against the CPU oracle at this size (unfused: 0.5 GB of grouped tensors for the one pair), strict and tensor-core builds
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    ref = O.cmflow_forward(sd, *inp[:4])
    print("N=4096 fp32", check_outputs(out, ref))
    net.set_precision("fp16x3")
    print("N=4096 fp16x3", check_outputs(run(net, inp), ref))
    # two pairs in one chunk == each pair alone (row-sliced global max, per-pair scales)
    inp2 = make_pairs(2, 4096, seed=3, dense=True)
    both = run(net, inp2)
    one = run(net, tuple(t[:1].contiguous() for t in inp2))
    assert torch.equal(both["sf_agg"][:1], one["sf_agg"]) and torch.equal(both["pre_trans"][:1], one["pre_trans"])


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
@pytest.mark.parametrize("B,N", [(2, 8), (3, 33), (1, 129)])
def test_tiny_and_ragged_clouds_match_oracle(B, N, precision):
    """Smallest legal cloud (N = 8 = the k of knn_point), N below the largest nsample, N not a multiple of any tile: whole forward against the
    CPU oracle.  (Tiles of the tensor-core kernels are mostly padding here; every row / column guard is exercised.)"""
    sd = synthetic_state_dict(0)
    net = CMFlow(Args()); net.load_state_dict(sd); net = net.to(DEV); net.set_precision(precision)
    inp = make_pairs(B, N, seed=21 + N)
    out = run(net, inp)
    ref = O.cmflow_forward(sd, *inp[:4], return_intermediates=True)
    assert knn_sets_equal(net.tap("knn12", (B, N, 8), torch.int32).cpu(), ref["knn12"].long().sort(-1)[0])
    print(B, N, precision, check_outputs(out, ref))


def test_fused_setconv1_equals_layerwise(golden_dir, monkeypatch):
    """The fused gather+MLP+max kernel of set-conv #1 against the GEMM-per-layer path (same fp32 FMAs, bias added first
    instead of last): agreement to fp32 rounding."""
    gold = load_golden(golden_dir, "cmflow_synth_b3_n200.pt")
    inp = case_inputs(gold["meta"])
    net, sd = build(gold["meta"], golden_dir)
    out = run(net, inp)
    f1 = net.tap("E", (3, 200, 800)).cpu()[..., :256]
    f2 = net.tap("f2", (3, 200, 256)).cpu()
    monkeypatch.setenv("CMF_FUSED_SC1", "0")
    net0, _ = build(gold["meta"], golden_dir)
    out0 = run(net0, inp)
    assert rel_err(f1, net0.tap("E", (3, 200, 800)).cpu()[..., :256]) < 2e-6
    assert rel_err(f2, net0.tap("f2", (3, 200, 256)).cpu()) < 2e-6
    assert rel_err(out["sf_agg"], out0["sf_agg"]) < 2e-5


@pytest.mark.parametrize("B,N,n_unique", [(3, 256, 120), (2, 200, 200), (2, 33, 20), (1, 1000, 400), (2, 129, 129), (1, 2304, 1500), (1, 4096, 4096)])
@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_thread_per_query_search_equals_warp_kernels(monkeypatch, B, N, n_unique, precision):
    """Small clouds run the neighbour search with one thread per query, dense ones with the warp-cooperative kernels: same distance
    functions, same ordering rules -> the same index tables bit for bit, ties included (duplicate-padded clouds as the training loader
    makes them, dataset/vod.py:102-110), and the same forward."""
    from cmflow_b200.synth import make_padded_pairs
    net = CMFlow(Args()); net.load_state_dict(synthetic_state_dict(0)); net = net.to(DEV)
    net.set_precision(precision)
    inp = make_padded_pairs(B, N, n_unique, seed=3)[0] if n_unique < N else make_pairs(B, N, seed=3)

    def once():
        out = run(net, inp)
        for k, c in (("bq1", 60), ("bq2", 60), ("knn12", 8), ("knn11", 8)):
            out[k] = net.tap(k, (B, N, c), torch.int32).cpu()
        return out
    monkeypatch.setenv("CMF_SEARCH_THREAD", "1")          # (small batches take the warp kernels by default)
    small = once()
    monkeypatch.delenv("CMF_SEARCH_THREAD")
    monkeypatch.setenv("CMF_SEARCH_WARP", "1")
    warp = once()
    for k in ("bq1", "bq2", "knn12", "knn11", "sf_agg", "stat_cls", "pre_trans", "mask"):
        assert torch.equal(small[k], warp[k]), k
