"""weights.py packing + the engine's hoisting algebra, replayed on CPU against the oracle (no GPU)."""
import pytest
import torch

from cmflow_b200 import weights
from oracle import cmflow_oracle as O
from tests.helpers import case_inputs, case_weights, load_golden, rel_err
from tests.pipeline_emulator import emulate


@pytest.mark.parametrize("name,temporal", [("cmflow_synth_b2_n256.pt", False), ("cmflow_ckpt_b2_n256.pt", False),
                                           ("cmflow_t_synth_b2_n256.pt", True)])
def test_packed_pipeline_matches_oracle(golden_dir, name, temporal):
    gold = load_golden(golden_dir, name)
    sd = case_weights(gold["meta"], golden_dir)
    if sd is None:
        pytest.skip("reference checkpoint not available")
    pc1, pc2, ft1, ft2, _ = case_inputs(gold["meta"])
    blob = weights.pack(sd, temporal)
    em = emulate(blob, pc1, pc2, ft1, ft2, temporal=temporal, dtype=torch.float64)
    ref = O.cmflow_forward(sd, pc1, pc2, ft1, ft2, dtype=torch.float64, temporal=temporal, return_intermediates=True)
    # float32-rounded folded weights vs the un-folded fp64 oracle: agreement to ~1e-6 proves the algebra
    assert rel_err(em["f1"].permute(0, 2, 1), ref["f1"]) < 2e-6
    assert rel_err(em["cor"].permute(0, 2, 1), ref["cor"]) < 2e-6
    assert rel_err(em["prop"].permute(0, 2, 1), ref["prop"]) < 5e-6
    assert rel_err(em["flow"], ref["flow"]) < 5e-6
    assert (em["stat_cls"] - ref["stat_cls"]).abs().max() < 5e-6
    if temporal:
        assert rel_err(em["gfeat"], ref["gfeat"]) < 5e-6


def test_blob_layout_matches_engine_table():
    from cmflow_b200._lib import lib
    from cmflow_b200.synth import synthetic_state_dict
    for temporal in (False, True):
        blob = weights.pack(synthetic_state_dict(0, temporal), temporal)
        assert blob.size == lib().cmf_model_blob_floats(int(temporal))


@pytest.mark.parametrize("name", ["raflow_synth_b3_n256.pt", "raflow_ckpt_b3_n256.pt"])
def test_packed_raflow_pipeline_matches_oracle(golden_dir, name):
    """RaFlow's state_dict mapped onto the CMFlow blob (weights.raflow_as_cmflow: fd_layer.mse -> set-conv #2, fd_layer.fp -> flow head, zero
    motion head): the engine's stage sequence replayed on the packed blob in fp64 gives the oracle's initial flow `output` (raflow.py:47-77)."""
    gold = load_golden(golden_dir, name)
    sd = case_weights(gold["meta"], golden_dir)
    if sd is None:
        pytest.skip("reference checkpoint not available")
    pc1, pc2, ft1, ft2, _ = case_inputs(gold["meta"])
    blob = weights.pack(sd, raflow=True)
    em = emulate(blob, pc1, pc2, ft1, ft2, dtype=torch.float64)
    ref = O.raflow_forward(sd, pc1, pc2, ft1, ft2, gold["interval"], dtype=torch.float64, return_intermediates=True)
    assert rel_err(em["f1"].permute(0, 2, 1), ref["f1"]) < 2e-6
    assert rel_err(em["cor"].permute(0, 2, 1), ref["cor"]) < 2e-6
    assert rel_err(em["prop"].permute(0, 2, 1), ref["prop"]) < 5e-6
    assert rel_err(em["flow"], ref["output"]) < 5e-6
    assert rel_err(em["flow"].float(), gold["output"]) < 1e-4          # and the unmodified reference's output
