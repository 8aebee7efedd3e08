"""Parity with "the reference's own lib/src CUDA path" (BASELINE.json north_star), and the drop-in levels of INTEGRATION.md, on a GPU.

The reference here is the real thing: its UNMODIFIED Python (models/*.py, utils/model_utils/radarflow_util.py, lib/pointnet2_utils.py,
main_util.py -- staged git-ignored under oracle/_ref/py by __graft_entry__.build()) over its own lib/src/*.cu compiled for sm_100a
(oracle/_ref/libpointnet2_ref.so), with cuDNN / cuBLAS fp32 (TF32 off) for everything it delegates to torch.

  1. whole forward: cmflow_b200's engine vs that build -- neighbour indices bit-exact, flow / transform / scores <= 1e-4;
  2. Level 1 drop-in: the reference's lib/pointnet2_utils.py and whole model running on cmflow_b200's operator module
     (`pointnet2_cuda` from cmflow_b200/shim) -- bit-identical to the same Python on the reference's kernels;
  3. Level 3 drop-in: the reference's evaluation loop (main_util.py:106-203, eval_one_epoch, one pair per call, un-resampled clouds)
     driving cmflow_b200.cmflow.CMFlow in place of models/cmflow.py::CMFlow -- same metrics.
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

from cmflow_b200.cmflow import CMFlow, CMFlow_T, RaFlow   # noqa: E402
from cmflow_b200.synth import make_pairs, raflow_state_dict, synthetic_state_dict   # noqa: E402
from oracle import ref_model as RM   # noqa: E402
from tests.helpers import case_weights, check_outputs, check_raflow_outputs, knn_sets_equal, load_golden, rel_err   # noqa: E402

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_ref():
    if not RM.available("cuda"):
        pytest.skip("oracle/_ref (reference kernels + staged reference Python) not built: run __graft_entry__.build() where /root/reference exists")
    RM.strict_fp32()


def ours(cls, sd, precision):
    net = cls(RM.Args()); net.load_state_dict(sd, strict=True); net = net.to(DEV); net.set_precision(precision)
    return net


def shim_module():
    """`import pointnet2_cuda` with cmflow_b200/shim on sys.path, as INTEGRATION.md Level 1 prescribes."""
    shim = os.path.join(ROOT, "cmflow_b200", "shim")
    sys.modules.pop("pointnet2_cuda", None)
    sys.path.insert(0, shim)
    try:
        import pointnet2_cuda
    finally:
        sys.path.remove(shim)
    assert os.path.dirname(os.path.abspath(pointnet2_cuda.__file__)) == shim
    return pointnet2_cuda


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
@pytest.mark.parametrize("case", ["cmflow_synth_b2_n256.pt", "cmflow_ckpt_b2_n256.pt", "cmflow_synth_b3_n200.pt"])
def test_cmflow_forward_matches_reference_cuda_build(golden_dir, case, precision):
    _need_ref()
    meta = load_golden(golden_dir, case)["meta"]
    sd = case_weights(meta, golden_dir)
    if sd is None:
        pytest.skip("reference checkpoint not available")
    ns = RM.load("cuda")
    ref_net = RM.build_model(ns, "cmflow", sd)
    B, N = meta["B"], meta["N"]
    inp = [t.to(DEV) for t in make_pairs(B, N, seed=meta["data_seed"])[:4]]
    with torch.no_grad():
        sf, cls, T, mask = ref_net(*inp, None, "test")
        x1, x2 = inp[0].transpose(2, 1).contiguous(), inp[1].transpose(2, 1).contiguous()
        ref_bq = torch.cat([ns.pointnet2_utils.ball_query(r, k, x1, x1) for r, k in ((2.0, 4), (4.0, 8), (8.0, 16), (16.0, 32))], -1)
        ref_knn12 = ns.radarflow_util.knn_point(8, x2, x1).sort(-1)[0]
        ref_knn11 = ns.radarflow_util.knn_point(8, x1, x1).sort(-1)[0]
    ref = {"sf_agg": sf.cpu(), "stat_cls": cls.cpu(), "pre_trans": T.cpu(), "mask": mask.cpu()}
    net = ours(CMFlow, sd, precision)
    with torch.no_grad():
        o = net(*inp, None, "test")
    out = {"sf_agg": o[0].cpu(), "stat_cls": o[1].cpu(), "pre_trans": o[2].cpu(), "mask": o[3].cpu()}
    assert torch.equal(net.tap("bq1", (B, N, 60), torch.int32), ref_bq)                      # ball-query tables: bit-exact, order included
    assert knn_sets_equal(net.tap("knn12", (B, N, 8), torch.int32).cpu(), ref_knn12.cpu())  # topk(sorted=False): sets
    assert knn_sets_equal(net.tap("knn11", (B, N, 8), torch.int32).cpu(), ref_knn11.cpu())
    print(case, precision, "vs reference CUDA build:", check_outputs(out, ref))


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_temporal_and_raflow_match_reference_cuda_build(golden_dir, precision):
    _need_ref()
    ns = RM.load("cuda")
    # CMFlow-T: three frames, GRU state carried
    meta = load_golden(golden_dir, "cmflow_t_synth_b2_n256.pt")["meta"]
    sd = synthetic_state_dict(meta["weight_seed"], temporal=True)
    ref_net, net = RM.build_model(ns, "cmflow_t", sd), ours(CMFlow_T, sd, precision)
    inp = [t.to(DEV) for t in make_pairs(meta["B"], meta["N"], seed=meta["data_seed"])[:4]]
    g_ref, g = None, None
    for step in range(3):
        with torch.no_grad():
            r = ref_net(*inp, None, "test", g_ref)
            o = net(*inp, None, "test", g)
        errs = check_outputs({"sf_agg": o[0].cpu(), "stat_cls": o[1].cpu(), "pre_trans": o[2].cpu(), "mask": o[3].cpu()},
                             {"sf_agg": r[0].cpu(), "stat_cls": r[1].cpu(), "pre_trans": r[2].cpu(), "mask": r[3].cpu()})
        assert rel_err(o[4].cpu(), r[4].cpu()) <= 1e-4
        print("cmflow_t step", step, precision, errs)
        g_ref, g = r[4], o[4]
    # RaFlow: both refinement branches (golden seeds keep every residual off the threshold)
    gold = load_golden(golden_dir, "raflow_synth_b3_n256.pt")
    meta = gold["meta"]
    sdr = raflow_state_dict(meta["weight_seed"])
    ref_net, net = RM.build_model(ns, "raflow", sdr), ours(RaFlow, sdr, precision)
    inp = [t.to(DEV) for t in make_pairs(meta["B"], meta["N"], seed=meta["data_seed"])[:4]]
    itv = gold["interval"].to(DEV)
    with torch.no_grad():
        r = ref_net(*inp, itv)
        o = net(*inp, itv)
    keys = ("output", "sf_agg", "pre_trans", "mask_s")
    print("raflow", precision, check_raflow_outputs({k: v.cpu() for k, v in zip(keys, o)}, {k: v.cpu() for k, v in zip(keys, r)}))


def test_level1_reference_pointnet2_utils_over_our_operator_module():
    """lib/pointnet2_utils.py, unmodified, once over the reference's kernels and once over cmflow_b200/shim: every operator bit-identical,
    including the backward of grouping / gathering / interpolation through the reference's own autograd Functions."""
    _need_ref()
    ref = RM.load("cuda").pointnet2_utils
    mine = RM.load("cuda", pointnet2_module=shim_module()).pointnet2_utils
    assert ref is not mine and mine.pointnet2.__name__ == "pointnet2_cuda"
    g = torch.Generator().manual_seed(3)
    B, N, M, C = 3, 300, 77, 19
    xyz = (torch.rand(B, N, 3, generator=g) * torch.tensor([20.0, 10.0, 2.0])).to(DEV)
    new_xyz = xyz[:, :M].contiguous()
    feats = torch.randn(B, C, N, generator=g).to(DEV)
    for P in (ref, mine):
        P.out = {}
        P.out["fps"] = P.furthest_point_sample(xyz, 40)
        P.out["gather"] = P.gather_operation(feats, P.out["fps"])
        P.out["bq"] = P.ball_query(3.0, 16, xyz, new_xyz)
        P.out["group"] = P.grouping_operation(feats, P.out["bq"])
        d, i = P.knn(8, new_xyz, xyz)
        P.out["knn_d"], P.out["knn_i"] = d, i
        d3, i3 = P.three_nn(xyz, new_xyz)
        P.out["nn3_d"], P.out["nn3_i"] = d3, i3
        w = 1.0 / (d3 + 1e-8); w = w / w.sum(2, keepdim=True)
        P.out["interp"] = P.three_interpolate(feats[:, :, :M].contiguous(), i3, w)
        P.out["qg"] = P.QueryAndGroup(3.0, 16)(xyz, new_xyz, feats)
        f = feats.clone().requires_grad_(True)
        P.grouping_operation(f, P.out["bq"]).square().sum().backward()
        P.out["group_grad"] = f.grad.clone()
        f = feats.clone().requires_grad_(True)
        P.gather_operation(f, P.out["fps"]).square().sum().backward()
        P.out["gather_grad"] = f.grad.clone()
        f = feats[:, :, :M].clone().requires_grad_(True)
        P.three_interpolate(f, i3, w).square().sum().backward()
        P.out["interp_grad"] = f.grad.clone()
    for k in ref.out:
        if k.endswith("_grad"):            # atomicAdd order is free in both builds: equal up to fp32 summation order
            assert rel_err(mine.out[k], ref.out[k], per_pair=False) <= 1e-6, k
        else:
            assert torch.equal(mine.out[k], ref.out[k]), k


def test_level1_reference_model_over_our_operator_module(golden_dir):
    """The reference's whole CMFlow (its Python, its torch ops) with only `pointnet2_cuda` swapped for cmflow_b200/shim: the operators are
    integer / copy work, so the forward is bit-identical to the reference on its own kernels."""
    _need_ref()
    meta = load_golden(golden_dir, "cmflow_synth_b2_n256.pt")["meta"]
    sd = synthetic_state_dict(meta["weight_seed"])
    inp = [t.to(DEV) for t in make_pairs(meta["B"], meta["N"], seed=meta["data_seed"])[:4]]
    outs = []
    for mod in (None, shim_module()):
        net = RM.build_model(RM.load("cuda", pointnet2_module=mod), "cmflow", sd)
        with torch.no_grad():
            outs.append([t.clone() for t in net(*inp, None, "test")])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


class _Loader(list):
    batch_size = 1


class _Text:
    def cprint(self, s):
        print(s)


def test_level3_reference_eval_loop_drives_our_model(golden_dir):
    """eval_one_epoch of the reference's main_util.py (unmodified), one pair per call as main.py:203 builds its test loader, over
    un-resampled clouds whose size changes every frame (N1 != N2): cmflow_b200.cmflow.CMFlow in the place of models.cmflow.CMFlow."""
    _need_ref()
    gold = load_golden(golden_dir, "real_radar_ckpt_n1n2.pt")
    sd = case_weights(gold["meta"], golden_dir)
    if sd is None:
        sd = synthetic_state_dict(0)
    ns = RM.load_eval_loop("cuda")
    g = torch.Generator().manual_seed(8)
    frames = []
    clouds = [(fr["pc1"], fr["pc2"], fr["ft1"], fr["ft2"]) for fr in gold["frames"]]
    for k, (n1, n2) in enumerate(((190, 231), (305, 288), (256, 256))):
        a, b = make_pairs(1, n1, seed=40 + k), make_pairs(1, n2, seed=60 + k)
        clouds.append((a[0], b[1], a[2], b[3]))
    for pc1, pc2, ft1, ft2 in clouds:
        n1 = pc1.shape[2]
        yaw = torch.tensor(0.01)
        T = torch.eye(4).unsqueeze(0)
        T[0, 0, 0], T[0, 0, 1], T[0, 1, 0], T[0, 1, 1], T[0, 0, 3] = yaw.cos(), -yaw.sin(), yaw.sin(), yaw.cos(), 0.8
        gt = ((T[:, :3, :3] @ pc1 + T[:, :3, 3:]) - pc1).transpose(2, 1).contiguous() + 0.02 * torch.randn(1, n1, 3, generator=g)
        mask = (torch.rand(1, n1, generator=g) > 0.3).float()
        z = torch.zeros(1, n1)
        # the loader's tuple (dataset/vod.py:117; main_util.py:121): clouds are (B,N,3) there and transposed by the loop
        frames.append((pc1.transpose(2, 1).contiguous(), pc2.transpose(2, 1).contiguous(), ft1.transpose(2, 1).contiguous(),
                       ft2.transpose(2, 1).contiguous(), T.clone(), gt, mask, torch.tensor([0.1]), z, z, torch.zeros(1, n1, 2)))

    class A:
        model = "cmflow"
        save_res = False
        vis = False
        radar_res = {"r_res": 0.2, "theta_res": 1.5 * 3.141592653589793 / 180, "phi_res": 1.5 * 3.141592653589793 / 180}   # dataset/vod.py:21-23

    ref_net = RM.build_model(ns, "cmflow", sd)
    res_ref = ns.main_util.eval_one_epoch(A(), ref_net, _Loader(frames), _Text())
    for precision in ("fp32", "fp16x3"):
        net = ours(CMFlow, sd, precision)
        res = ns.main_util.eval_one_epoch(A(), net, _Loader(frames), _Text())
        for d_ours, d_ref in zip(res[:3], res_ref[:3]):
            for k in d_ref:
                a, b = float(d_ours[k]), float(d_ref[k])
                # accuracy / segmentation scores count points against thresholds: allow one borderline point over the ~1500 evaluated
                tol = 2e-3 if k in ("sas", "ras", "accs", "accr", "acc", "miou", "sen") else 1e-4 * max(abs(b), 1.0)
                assert abs(a - b) <= tol, (precision, k, a, b)
        assert rel_err(res[4][:, :3].cpu(), res_ref[4][:, :3].cpu()) <= 1e-4
        print(precision, "eval loop metrics:", {k: float(v) for k, v in res[0].items()}, {k: float(v) for k, v in res[2].items()})
