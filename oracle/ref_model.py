"""The UNMODIFIED reference Python model, importable where /root/reference does not exist (TEST INFRASTRUCTURE; GPU box).

`stage()` (run in the build container, from __graft_entry__.build) copies the reference's Python packages that the hot path imports
-- models/, utils/, lib/*.py, losses/ -- byte for byte into oracle/_ref/py/ (git-ignored like the rest of oracle/_ref, so nothing of
the reference enters the history; it travels to the GPU box with the repo snapshot exactly like oracle/_ref/libpointnet2_ref.so).

`load(device)` imports them from there with
  * stubs for third-party packages the reference imports but the forward never touches (open3d, ujson, h5py, cv2, matplotlib):
    utils/__init__.py:4 -> vis_util.py:6-13;
  * a module named `pointnet2_cuda` (imported at lib/pointnet2_utils.py:7) whose ten functions have the positional signatures of
    lib/src/pointnet2_api.cpp:11-24:
        device="cuda": backed by oracle/_ref/libpointnet2_ref.so = the reference's own lib/src/*.cu compiled unmodified for sm_100a.
                       This is "the reference's own lib/src CUDA build" of BASELINE.json's north_star: reference Python + reference
                       kernels, cuDNN/cuBLAS for everything the reference delegates to torch, TF32 switched off.
        device="cpu":  backed by oracle/pointops_oracle.c, with `.cuda()` a no-op and torch.cuda.FloatTensor / IntTensor aliased to
                       the CPU tensor types (lib/pointnet2_utils.py allocates its outputs with them) -- the reference's only way to run
                       without a GPU, used by tests/golden/make_golden.py and by bench.py's --impl reference arm.
No file of the reference is edited.  Never imported by cmflow_b200/.
"""
import os
import shutil
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref", "py")
REF = "/root/reference"
PACKAGES = ("models", "utils", "losses")
LIB_FILES = ("pointnet2_utils.py", "pointnet2_modules.py", "pytorch_utils.py")
TOP_FILES = ("main_util.py", "clip_util.py")        # the evaluation loops (eval_one_epoch, main_util.py:106-203; eval_one_epoch_seq, clip_util.py)


def stage(ref_root=REF):
    """Copy the reference's Python packages into oracle/_ref/py (no-op when the reference tree is absent)."""
    if not os.path.isdir(ref_root):
        return STAGED if os.path.isdir(STAGED) else None
    os.makedirs(STAGED, exist_ok=True)
    for pkg in PACKAGES:
        dst = os.path.join(STAGED, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(ref_root, pkg), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(STAGED, "lib"), exist_ok=True)
    for f in LIB_FILES:
        shutil.copyfile(os.path.join(ref_root, "lib", f), os.path.join(STAGED, "lib", f))
    for f in TOP_FILES:
        shutil.copyfile(os.path.join(ref_root, f), os.path.join(STAGED, f))
    return STAGED


def root():
    """Directory to put on sys.path: the live reference tree when present, else the staged copy."""
    if os.path.isdir(os.path.join(REF, "models")):
        return REF
    if os.path.isdir(os.path.join(STAGED, "models")):
        return STAGED
    return None


def available(device="cuda"):
    if root() is None:
        return False
    if device == "cuda":
        from . import refcuda
        return refcuda.available() and torch.cuda.is_available()
    return True


def _stub_third_party():
    for name in ("open3d", "ujson", "h5py", "cv2", "matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.mlab"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.setdefault("MultipleLocator", object)
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].ticker = sys.modules["matplotlib.ticker"]
    sys.modules["matplotlib"].mlab = sys.modules["matplotlib.mlab"]


def refcuda_pointnet2_module():
    """`pointnet2_cuda` over the reference's own compiled kernels: same positional signatures, caller-allocated outputs
    (lib/src/pointnet2_api.cpp:11-24; wrappers in ball_query.cpp:14-25, group_points.cpp:12-36, sampling.cpp:11-45, interpolate.cpp:14-71)."""
    import ctypes

    from . import refcuda
    L = refcuda.lib()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    s = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    f = ctypes.c_float
    m = types.ModuleType("pointnet2_cuda")

    def ball_query_wrapper(b, n, mm, radius, nsample, new_xyz, xyz, idx):
        L.ref_ball_query(b, n, mm, f(radius), nsample, p(new_xyz), p(xyz), p(idx), s()); return 1

    def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
        L.ref_group_points(b, c, n, npoints, nsample, p(points), p(idx), p(out), s()); return 1

    def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
        L.ref_group_points_grad(b, c, n, npoints, nsample, p(grad_out), p(idx), p(grad_points), s()); return 1

    def gather_points_wrapper(b, c, n, npoints, points, idx, out):
        L.ref_gather_points(b, c, n, npoints, p(points), p(idx), p(out), s()); return 1

    def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
        L.ref_gather_points_grad(b, c, n, npoints, p(grad_out), p(idx), p(grad_points), s()); return 1

    def furthest_point_sampling_wrapper(b, n, mm, points, temp, idx):
        L.ref_furthest_point_sampling(b, n, mm, p(points), p(temp), p(idx), s()); return 1

    def knn_wrapper(b, n, mm, k, unknown, known, dist2, idx):
        L.ref_knn(b, n, mm, k, p(unknown), p(known), p(dist2), p(idx), s())

    def three_nn_wrapper(b, n, mm, unknown, known, dist2, idx):
        L.ref_three_nn(b, n, mm, p(unknown), p(known), p(dist2), p(idx), s())

    def three_interpolate_wrapper(b, c, mm, n, points, idx, weight, out):
        L.ref_three_interpolate(b, c, mm, n, p(points), p(idx), p(weight), p(out), s())

    def three_interpolate_grad_wrapper(b, c, n, mm, grad_out, idx, weight, grad_points):
        L.ref_three_interpolate_grad(b, c, n, mm, p(grad_out), p(idx), p(weight), p(grad_points), s())

    for fn in (ball_query_wrapper, group_points_wrapper, group_points_grad_wrapper, gather_points_wrapper, gather_points_grad_wrapper,
               furthest_point_sampling_wrapper, knn_wrapper, three_nn_wrapper, three_interpolate_wrapper, three_interpolate_grad_wrapper):
        setattr(m, fn.__name__, fn)
    return m


_loaded = {}


def _purge_reference_modules():
    for name in list(sys.modules):
        if name in ("models", "utils", "lib", "losses", "pointnet2_cuda", "main_util", "clip_util") or name.split(".")[0] in ("models", "utils", "lib", "losses"):
            del sys.modules[name]


def load(device="cuda", pointnet2_module=None):
    """Import the reference model classes.  Returns a namespace with CMFlow, CMFlow_T, RaFlow, radarflow_util, pointnet2_utils
    (the reference's modules, untouched).  `pointnet2_module` overrides what `pointnet2_cuda` resolves to (the Level-1 drop-in test
    passes cmflow_b200's own operator module here)."""
    key = (device, id(pointnet2_module))
    if key in _loaded:
        return _loaded[key]
    r = root()
    if r is None:
        raise RuntimeError("reference Python is neither at /root/reference nor staged under oracle/_ref/py (run __graft_entry__.build() in the build container)")
    _purge_reference_modules()
    _stub_third_party()
    if device == "cpu":
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.FloatTensor = torch.FloatTensor
        torch.cuda.IntTensor = torch.IntTensor
        from . import pointops
        mod = pointnet2_module or pointops.as_pointnet2_module()
    else:
        try:                                     # legacy typed constructors the reference allocates its outputs with (pointnet2_utils.py:93,200,246)
            torch.cuda.FloatTensor(1, 1)
        except Exception:                        # a torch without them: same semantics (uninitialised CUDA tensor of that shape)
            torch.cuda.FloatTensor = lambda *sz: torch.empty(*sz, dtype=torch.float32, device="cuda")
            torch.cuda.IntTensor = lambda *sz: torch.empty(*sz, dtype=torch.int32, device="cuda")
        mod = pointnet2_module or refcuda_pointnet2_module()
    sys.modules["pointnet2_cuda"] = mod
    if r not in sys.path:
        sys.path.insert(0, r)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # the reference's own SyntaxWarnings (models/model.py:30)
        from lib import pointnet2_utils
        from models.cmflow import CMFlow
        from models.cmflow_t import CMFlow_T
        from models.raflow import RaFlow
        from utils.model_utils import radarflow_util
    ns = types.SimpleNamespace(CMFlow=CMFlow, CMFlow_T=CMFlow_T, RaFlow=RaFlow, radarflow_util=radarflow_util,
                               pointnet2_utils=pointnet2_utils, root=r, device=device)
    _loaded[key] = ns
    return ns


def load_eval_loop(device="cuda", pointnet2_module=None):
    """The reference's own evaluation loop module (main_util.py: eval_one_epoch, :106-203), imported unmodified.  It does
    `from time import clock` (main_util.py:7; removed in Python 3.8): time.clock is aliased to time.perf_counter for the import."""
    import time
    ns = load(device, pointnet2_module)
    if not hasattr(time, "clock"):
        time.clock = time.perf_counter
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import main_util
    ns.main_util = main_util
    return ns


def strict_fp32():
    """TF32 off for matmul and cuDNN: the reference predates TF32 defaults (PyTorch 1.7, src/GETTING_STARTED.md:40) and the parity bar is fp32."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.set_float32_matmul_precision("highest")
    except Exception:
        pass


class Args:
    num_points = 256
    stat_thres = 0.5
    rigid_thres = 0.15


def build_model(ns, kind, state_dict, device="cuda"):
    cls = {"cmflow": ns.CMFlow, "cmflow_t": ns.CMFlow_T, "raflow": ns.RaFlow}[kind]
    net = cls(Args())
    missing = net.load_state_dict(state_dict, strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys, missing
    net = net.eval()
    return net.cuda() if device == "cuda" else net
