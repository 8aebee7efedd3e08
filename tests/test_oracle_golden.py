"""The CPU oracle (oracle/) pinned against golden vectors produced by the UNMODIFIED reference Python
(tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import cmflow_oracle as O
from oracle import pointops as P
from tests.helpers import case_inputs, case_weights, check_outputs, check_raflow_outputs, knn_sets_equal, load_golden, rel_err

CASES = ["cmflow_synth_b2_n256.pt", "cmflow_synth_w1_b2_n256.pt", "cmflow_synth_b3_n200.pt", "cmflow_synth_b2_n40.pt",
         "cmflow_ckpt_b2_n256.pt"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_forward_matches_reference(golden_dir, name):
    gold = load_golden(golden_dir, name)
    sd = case_weights(gold["meta"], golden_dir)
    if sd is None:
        pytest.skip("reference checkpoint not available")
    pc1, pc2, ft1, ft2, _ = case_inputs(gold["meta"])
    out = O.cmflow_forward(sd, pc1, pc2, ft1, ft2, return_intermediates=True)
    # integer work: exact
    assert knn_sets_equal(out["knn12"], gold["knn12"])
    assert knn_sets_equal(out["knn11"], gold["knn11"])
    # stage boundaries (every 4th point of pair 0) and final outputs within the north-star bar
    for key, val in (("f1_sub", out["f1"]), ("f2_sub", out["f2"]), ("cor_sub", out["cor"]), ("prop_sub", out["prop"])):
        assert rel_err(val[0, :, ::4], gold[key], per_pair=False) <= 1e-4, key
    errs = check_outputs(out, gold)
    print(name, errs)


def test_oracle_temporal_matches_reference(golden_dir):
    gold = load_golden(golden_dir, "cmflow_t_synth_b2_n256.pt")
    sd = case_weights(gold["meta"], golden_dir)
    pc1, pc2, ft1, ft2, _ = case_inputs(gold["meta"])
    g = None
    for step in gold["steps"]:
        out = O.cmflow_forward(sd, pc1, pc2, ft1, ft2, temporal=True, gfeat_prev=g)
        check_outputs(out, step)
        assert rel_err(out["gfeat"], step["gfeat"]) <= 1e-4
        g = out["gfeat"]


def test_oracle_kabsch_matches_reference(golden_dir):
    gold = load_golden(golden_dir, "kabsch_n128.pt")
    T, _ = O.weighted_kabsch(gold["A"], gold["B"], gold["W"])
    assert rel_err(T[:, :3], gold["T"][:, :3]) <= 1e-4
    # reflected cloud (case 3) must reproduce the reference's row-2 flip: R = diag(1,1,-1) V U^T, det(R) = +1
    assert torch.linalg.det(T[3, :3, :3]) > 0.99
    T64, _ = O.weighted_kabsch(gold["A"].double(), gold["B"].double(), gold["W"].double())
    assert rel_err(T64[:, :3].float(), gold["T"][:, :3]) <= 1e-4


def test_oracle_fp64_truth_close_to_fp32(golden_dir):
    gold = load_golden(golden_dir, "cmflow_synth_b2_n256.pt")
    sd = case_weights(gold["meta"], golden_dir)
    pc1, pc2, ft1, ft2, _ = case_inputs(gold["meta"])
    out = O.cmflow_forward(sd, pc1, pc2, ft1, ft2, dtype=torch.float64)
    out = {k: (v.float() if v.dtype == torch.float64 else v) for k, v in out.items()}
    check_outputs(out, gold)


@pytest.mark.parametrize("name", ["raflow_synth_b3_n256.pt", "raflow_ckpt_b3_n256.pt"])
def test_oracle_raflow_matches_reference(golden_dir, name):
    """oracle.raflow_forward against the unmodified models/raflow.py (both SFR branches: > / < 25 % rigid inliers)."""
    gold = load_golden(golden_dir, name)
    sd = case_weights(gold["meta"], golden_dir)
    if sd is None:
        pytest.skip("reference checkpoint not available")
    pc1, pc2, ft1, ft2, _ = case_inputs(gold["meta"])
    out = O.raflow_forward(sd, pc1, pc2, ft1, ft2, gold["interval"], return_intermediates=True)
    for key, val in (("f1_sub", out["f1"]), ("f2_sub", out["f2"]), ("cor_sub", out["cor"]), ("prop_sub", out["prop"])):
        assert rel_err(val[0, :, ::4], gold[key], per_pair=False) <= 1e-4, key
    frac = gold["mask_s"].float().mean(1)
    assert (frac > 0.25).any() and (frac < 0.25).any()
    print(name, check_raflow_outputs(out, gold))
    out64 = O.raflow_forward(sd, pc1, pc2, ft1, ft2, gold["interval"], dtype=torch.float64)
    check_raflow_outputs({k: (v.float() if v.dtype == torch.float64 else v) for k, v in out64.items()}, gold)


# ---- round-2 fixtures (tests/golden/make_golden.py --round2-only) ------------------------------------------------------------------
def uid_sets(idx, mapping):
    """(B,N,k) neighbour indices -> sorted unique ids of the points they refer to (duplicate-padded clouds: which copy of a duplicated
    point enters a neighbour set is implementation-defined under torch.topk, the point itself is not)."""
    B, N, k = idx.shape
    return torch.gather(mapping.long(), 1, idx.long().flatten(1)).view(B, N, k).sort(-1)[0]


def test_oracle_duplicate_padded_clouds_match_reference(golden_dir):
    """Clouds padded to num_points by duplicate sampling as the training loader does (dataset/vod.py:102-110): exact distance ties."""
    from cmflow_b200.synth import make_padded_pairs
    gold = load_golden(golden_dir, "cmflow_synth_padded_b2_n256.pt")
    meta = gold["meta"]
    sd = case_weights(meta, golden_dir)
    (pc1, pc2, ft1, ft2, _), maps, _ = make_padded_pairs(meta["B"], meta["N"], meta["n_unique"], meta["data_seed"])
    assert torch.equal(maps[0].to(torch.int16), gold["map1"]) and torch.equal(maps[1].to(torch.int16), gold["map2"])
    out = O.cmflow_forward(sd, pc1, pc2, ft1, ft2, return_intermediates=True)
    assert torch.equal(uid_sets(out["knn12"], maps[1]), gold["knn12_uid"].long())
    assert torch.equal(uid_sets(out["knn11"], maps[0]), gold["knn11_uid"].long())
    for key, val in (("f1_sub", out["f1"]), ("f2_sub", out["f2"]), ("cor_sub", out["cor"]), ("prop_sub", out["prop"])):
        assert rel_err(val[0, :, ::4], gold[key], per_pair=False) <= 1e-4, key
    print(check_outputs(out, gold))


def test_oracle_train_mode_labels_match_reference(golden_dir):
    """mode='train' with pseudo labels (models/cmflow.py:181-182): labels drive the Kabsch weights and the refinement mask."""
    gold = load_golden(golden_dir, "cmflow_synth_train_b2_n256.pt")
    sd = case_weights(gold["meta"], golden_dir)
    pc1, pc2, ft1, ft2, _ = case_inputs(gold["meta"])
    out = O.cmflow_forward(sd, pc1, pc2, ft1, ft2, label_m=gold["label_m"])
    assert torch.equal(out["mask"], gold["mask"]) and torch.equal(out["mask"], gold["label_m"] > 0.5)
    assert rel_err(out["pre_trans"][:, :3], gold["pre_trans"][:, :3]) <= 1e-4
    assert rel_err(out["sf_agg"], gold["sf_agg"]) <= 1e-4
    assert (out["stat_cls"] - gold["stat_cls"]).abs().max() <= 1e-4


def test_oracle_real_radar_frames_n1_ne_n2(golden_dir):
    """Real radar clouds of the reference's own test run, un-resampled (N1 != N2), one pair per call as main.py:203 evaluates."""
    gold = load_golden(golden_dir, "real_radar_ckpt_n1n2.pt")
    sd = case_weights(gold["meta"], golden_dir)
    sdr = case_weights({"weights": gold["meta"]["weights_raflow"], "model": "raflow"}, golden_dir)
    if sd is None or sdr is None:
        pytest.skip("reference checkpoints not available")
    for fr in gold["frames"]:
        assert fr["pc1"].shape[2] != fr["pc2"].shape[2]
        out = O.cmflow_forward(sd, fr["pc1"], fr["pc2"], fr["ft1"], fr["ft2"], return_intermediates=True)
        assert knn_sets_equal(out["knn12"], fr["cmflow"]["knn12"]) and knn_sets_equal(out["knn11"], fr["cmflow"]["knn11"])
        assert rel_err(out["prop"][0, :, ::4], fr["cmflow"]["prop_sub"], per_pair=False) <= 1e-4
        print(fr["source"], check_outputs(out, fr["cmflow"]))
        if fr["raflow"] is not None:
            outr = O.raflow_forward(sdr, fr["pc1"], fr["pc2"], fr["ft1"], fr["ft2"], fr["raflow"]["interval"])
            print(fr["source"], check_raflow_outputs(outr, fr["raflow"]))


def test_oracle_illconditioned_kabsch(golden_dir):
    """WeightedKabsch on near-planar / near-collinear clouds and on weights concentrated on 3-4 points: the fp32 torch.svd result of the
    reference against the oracle in fp32 and fp64 (R = V U^T is the polar factor of H: well defined as long as sigma_2 + sigma_3 > 0)."""
    gold = load_golden(golden_dir, "kabsch_illcond_n128.pt")
    T32, _ = O.weighted_kabsch(gold["A"], gold["B"], gold["W"])
    T64, H = O.weighted_kabsch(gold["A"].double(), gold["B"].double(), gold["W"].double())
    sv = torch.linalg.svdvals(H)
    print("sigma3/sigma1", (sv[:, 2] / sv[:, 0]).tolist())
    e32, e64 = rel_err(T32[:, :3], gold["T"][:, :3]), rel_err(T64[:, :3].float(), gold["T"][:, :3])
    print("oracle fp32 vs reference", e32, "oracle fp64 vs reference", e64)
    assert e32 <= 1e-4 and e64 <= 1e-4
