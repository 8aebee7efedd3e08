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
