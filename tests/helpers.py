"""Shared comparison helpers for the parity tests (tolerances are stated here, once).

FLOW_RTOL = 1e-4 is BASELINE.json north_star's bar for flow vectors and the rigid transform:
|a-b| <= FLOW_RTOL * max(|b|, scale) elementwise, where `scale` is the largest magnitude of the
reference tensor for that frame pair (so near-zero components are judged against the vector's size).
Integer / index outputs are compared exactly.
"""
import os

import torch

FLOW_RTOL = 1e-4
THRESH_GUARD = 1e-4      # points whose static score is this close to stat_thres may flip mask legitimately


def rel_err(a, b, per_pair=True):
    """max over elements of |a-b| / max(|b|, per-pair max|b|)."""
    a, b = a.double(), b.double()
    if per_pair and b.dim() >= 2:
        scale = b.abs().flatten(1).max(1)[0].clamp_min(1e-30).view(-1, *([1] * (b.dim() - 1)))
    else:
        scale = b.abs().max().clamp_min(1e-30)
    return ((a - b).abs() / scale).max().item()


def load_golden(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), map_location="cpu", weights_only=True)


def case_inputs(meta, **kw):
    from cmflow_b200.synth import make_pairs
    return make_pairs(meta["B"], meta["N"], seed=meta["data_seed"], **kw)


def stage_reference_weights(golden_dir, ref_root="/root/reference", parts=4):
    """Copy the reference's pretrained checkpoints next to the golden vectors (git-ignored; they travel to the GPU box with the
    repo snapshot, where /root/reference does not exist).  Split into a few files: large single files are not shipped."""
    out = os.path.join(golden_dir, "_weights")
    for exp in ("cmflow_cvpr", "raflow_cvpr"):
        src = os.path.join(ref_root, "checkpoints", exp, "models", "model.best.t7")
        if not os.path.exists(src) or os.path.exists(os.path.join(out, f"{exp}.part0.t7")):
            continue
        os.makedirs(out, exist_ok=True)
        sd = torch.load(src, map_location="cpu", weights_only=True)
        keys = list(sd)
        for i in range(parts):
            torch.save({k: sd[k].clone() for k in keys[i::parts]}, os.path.join(out, f"{exp}.part{i}.t7"))


def case_weights(meta, golden_dir):
    from cmflow_b200.synth import synthetic_state_dict
    if "weights" in meta:
        exp = os.path.basename(os.path.dirname(os.path.dirname(meta["weights"])))
        ref = os.path.join("/root/reference", meta["weights"])
        if os.path.exists(ref):
            return torch.load(ref, map_location="cpu", weights_only=True)
        parts = sorted(f for f in os.listdir(os.path.join(golden_dir, "_weights")) if f.startswith(exp + ".part")) \
            if os.path.isdir(os.path.join(golden_dir, "_weights")) else []
        if parts:
            merged = {}
            for f in parts:
                merged.update(torch.load(os.path.join(golden_dir, "_weights", f), map_location="cpu", weights_only=True))
            # restore the reference's key order (load_state_dict does not care, packing does not either; kept for tidiness)
            return merged
        return None
    if meta["model"] == "raflow":
        from cmflow_b200.synth import raflow_state_dict
        return raflow_state_dict(meta["weight_seed"])
    return synthetic_state_dict(meta["weight_seed"], temporal=(meta["model"] == "cmflow_t"))


def knn_sets_equal(idx, ref_sorted):
    """idx (B,N,k) any order vs reference sets sorted ascending (int16)."""
    return torch.equal(idx.long().sort(-1)[0], ref_sorted.long())


def check_outputs(out, gold, rtol=FLOW_RTOL, stat_thres=0.5):
    """Compare a forward result dict with a golden dict; returns dict of measured errors and asserts the bars."""
    cls_g = gold["stat_cls"]
    safe = ((cls_g - stat_thres).abs() > THRESH_GUARD).squeeze(1)            # (B,N)
    errs = {}
    errs["stat_cls"] = (out["stat_cls"].double() - cls_g.double()).abs().max().item()
    errs["pre_trans"] = rel_err(out["pre_trans"][:, :3, :], gold["pre_trans"][:, :3, :])
    m_ok = (out["mask"].bool() == gold["mask"].bool()) | ~safe
    errs["mask_mismatch"] = int((~m_ok).sum())
    a = torch.where(safe.unsqueeze(1), out["sf_agg"].double(), gold["sf_agg"].double())
    errs["sf_agg"] = rel_err(a, gold["sf_agg"])
    assert errs["mask_mismatch"] == 0, errs
    assert errs["stat_cls"] <= rtol, errs            # probabilities in [0,1]: absolute == relative to range
    assert errs["pre_trans"] <= rtol, errs
    assert errs["sf_agg"] <= rtol, errs
    return errs


def check_raflow_outputs(out, gold, rtol=FLOW_RTOL):
    """RaFlow.forward outputs (models/raflow.py:157-164): initial flow, aggregated flow, transform, rigid-inlier mask (exact: the
    golden seeds keep every point's |residual / vel| at least 2e-4 away from the threshold)."""
    errs = {"output": rel_err(out["output"], gold["output"]), "sf_agg": rel_err(out["sf_agg"], gold["sf_agg"]),
            "pre_trans": rel_err(out["pre_trans"][:, :3, :], gold["pre_trans"][:, :3, :]),
            "mask_mismatch": int((out["mask_s"].bool() != gold["mask_s"].bool()).sum())}
    assert errs["mask_mismatch"] == 0, errs
    assert errs["output"] <= rtol and errs["sf_agg"] <= rtol and errs["pre_trans"] <= rtol, errs
    return errs
