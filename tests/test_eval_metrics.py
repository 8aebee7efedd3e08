"""Evaluation metrics (utils/eval_util.py:42-117): the numpy oracle against golden values from the UNMODIFIED reference functions (CPU),
and the device reductions against both (GPU)."""
import numpy as np
import pytest
import torch

from oracle import eval_oracle as E
from tests.golden.make_golden import eval_case
from tests.helpers import load_golden


def _close(got, want, rtol):
    for k, v in want.items():
        assert abs(float(got[k]) - v) <= rtol * max(abs(v), 1e-12), (k, float(got[k]), v)


def test_eval_oracle_matches_reference(golden_dir):
    gold = load_golden(golden_dir, "eval_metrics.pt")
    pc1, pred, labels, mask, pred_m, T, pred_T = (t.numpy() for t in eval_case(gold["meta"]["seed"], gold["meta"]["B"], gold["meta"]["N"]))
    _close(E.eval_scene_flow(pc1, pred, labels, mask, gold["radar_res"]), gold["sf"], 1e-6)
    _close(E.eval_motion_seg(pred_m, mask), gold["seg"], 1e-12)
    _close(E.eval_trans_rpe(T, pred_T), gold["pose"], 1e-5)


@pytest.mark.gpu
def test_eval_kernels_match_reference(golden_dir):
    from cmflow_b200 import eval_util as G
    gold = load_golden(golden_dir, "eval_metrics.pt")
    pc1, pred, labels, mask, pred_m, T, pred_T = (t.cuda() for t in eval_case(gold["meta"]["seed"], gold["meta"]["B"], gold["meta"]["N"]))

    class A:
        radar_res = gold["radar_res"]

    _close(G.eval_scene_flow(pc1, pred, labels, mask, A()), gold["sf"], 1e-5)
    _close(G.eval_motion_seg(pred_m, mask), gold["seg"], 1e-12)
    _close(G.eval_trans_RPE(T, pred_T), gold["pose"], 1e-4)        # float32 transforms: |t| ~ 0.07 from differences of ~1 m terms


@pytest.mark.gpu
def test_eval_accumulator_is_the_reference_loop():
    """EvalAccumulator == sum over batches of batch_size * per-batch metric, / pairs (main_util.py:176-202), with ragged last batch,
    larger clouds and an all-static batch (mov_rne's 1e-6 guard)."""
    from cmflow_b200 import eval_util as G
    res = {"r_res": 0.2, "theta_res": 1.5 * np.pi / 180, "phi_res": 1.5 * np.pi / 180}
    acc = G.EvalAccumulator(res, "cuda")
    want = {k: 0.0 for k in G.SF_KEYS + G.SEG_KEYS + G.POSE_KEYS}
    pairs = 0
    for seed, B, N in ((3, 5, 300), (4, 2, 1024), (5, 1, 64)):
        pc1, pred, labels, mask, pred_m, T, pred_T = eval_case(seed, B, N)
        if seed == 5:
            mask = torch.ones_like(mask)
        acc.add(pc1.cuda(), pred.transpose(2, 1).contiguous().cuda(), labels.cuda(), mask.cuda(), pred_m.cuda(), T.cuda(), pred_T.cuda())
        ref = {**E.eval_scene_flow(pc1.numpy(), pred.numpy(), labels.numpy(), mask.numpy(), res), **E.eval_motion_seg(pred_m.numpy(), mask.numpy()),
               **E.eval_trans_rpe(T.numpy(), pred_T.numpy())}
        for k in want:
            want[k] += B * float(ref[k])
        pairs += B
    sf, seg, pose, n = acc.result(all_reduce=False)
    assert n == pairs
    _close({**sf, **seg, **pose}, {k: v / pairs for k, v in want.items()}, 1e-4)


def test_eval_has_no_cpu_path():
    from cmflow_b200 import eval_util as G
    from cmflow_b200._lib import CmfError
    with pytest.raises(CmfError):
        G.eval_motion_seg(torch.zeros(2, 4), torch.zeros(2, 4))
