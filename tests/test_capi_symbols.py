"""The C-ABI library loads on a CPU-only box and exports every entry point include/cmflow_b200.h declares.
No compute is called here (no GPU)."""
import ctypes
import os
import re

from cmflow_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "cmflow_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cmf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_table_agree():
    assert header_functions() == sorted(_lib.SIGNATURES)


def test_library_exports_every_symbol():
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(L, name), name


def test_version_and_error_plumbing_without_gpu():
    L = _lib.lib()
    assert b"sm_100a" in L.cmf_version()
    # invalid arguments are rejected before any CUDA call
    rc = L.cmf_knn(1, 4, 4, 500, None, None, None, None, None)
    assert rc == 1 and b"k must be" in L.cmf_last_error()
    rc = L.cmf_ball_query(1, 4, 4, 1.0, 2, None, None, None, None)
    assert rc == 1 and b"null pointer" in L.cmf_last_error()
    assert L.cmf_ball_query(0, 4, 4, 1.0, 2, None, None, None, None) == 0     # empty batch is a no-op
    assert _lib.watchdog_record() is None                                      # nothing has hung: the record is empty (and needs no GPU to read)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cmflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_raflow_mirror_has_the_reference_key_layout():
    """cmflow_b200.cmflow.RaFlow carries exactly the parameters / buffers of models/raflow.py (355 keys of checkpoints/raflow_cvpr)."""
    import torch
    from cmflow_b200 import weights
    from cmflow_b200.cmflow import RaFlow
    from cmflow_b200.synth import raflow_state_dict

    class A:
        num_points = 256
        rigid_thres = 0.15

    net = RaFlow(A())
    sd = raflow_state_dict(5)
    assert set(net.state_dict()) == set(sd) and len(sd) == 355
    net.load_state_dict(sd, strict=True)
    blob = weights.pack(net.state_dict(), raflow=True)
    assert blob.size == weights.pack(weights.raflow_as_cmflow(sd)).size
    ck = "/root/reference/checkpoints/raflow_cvpr/models/model.best.t7"
    import os
    if os.path.exists(ck):
        net.load_state_dict(torch.load(ck, map_location="cpu", weights_only=True), strict=True)
