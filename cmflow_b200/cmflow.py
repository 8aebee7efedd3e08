"""CMFlow / CMFlow_T with the reference's constructor, forward signature, return values and state_dict
key layout (models/cmflow.py:9-197, models/cmflow_t.py:10-211), running on the B200 engine.

    net = CMFlow(args).cuda(); net.load_state_dict(torch.load("checkpoints/cmflow_cvpr/models/model.best.t7"))
    sf_agg, stat_cls, pre_trans, mask = net(pc1, pc2, feature1, feature2, None, 'test')

Inference only (BatchNorm running statistics, no autograd through the engine); mode='train' with label_m feeds the
pseudo labels to the Kabsch head exactly as cmflow.py:181-182 does.  There is no CPU path.
"""
import ctypes
import warnings

import numpy as np
import torch
import torch.nn as nn

from . import weights as _weights
from ._lib import CmfError, check, dptr, lib, stream_ptr


class _SetConv(nn.Module):        # parameter container mirroring PointLocalFeature (radarflow_util.py:121-142)
    def __init__(self, in_channel, mlp, mlp2):
        super().__init__()
        self.mlp_convs, self.mlp_bns = nn.ModuleList(), nn.ModuleList()
        self.mlp2_convs, self.mlp2_bns = nn.ModuleList(), nn.ModuleList()
        last = in_channel + 3
        for co in mlp:
            self.mlp_convs.append(nn.Conv2d(last, co, 1, bias=False)); self.mlp_bns.append(nn.BatchNorm2d(co)); last = co
        for co in mlp2:
            self.mlp2_convs.append(nn.Conv2d(last, co, 1, bias=False)); self.mlp2_bns.append(nn.BatchNorm2d(co)); last = co


class _MultiScale(nn.Module):     # MultiScaleEncoder (radarflow_util.py:101-118)
    def __init__(self, in_channel, mlp, mlp2):
        super().__init__()
        self.ms_ls = nn.ModuleList([_SetConv(in_channel, mlp, mlp2) for _ in range(4)])


class _WeightNet(nn.Module):      # WeightNet (radarflow_util.py:288-305); BN modules exist but are unused (bn=False)
    def __init__(self):
        super().__init__()
        dims = [(3, 8), (8, 8), (8, 512)]
        self.mlp_convs = nn.ModuleList([nn.Conv2d(ci, co, 1) for ci, co in dims])
        self.mlp_bns = nn.ModuleList([nn.BatchNorm2d(co) for _, co in dims])


class _Correlator(nn.Module):     # FeatureCorrelator (radarflow_util.py:164-183)
    def __init__(self):
        super().__init__()
        self.mlp_convs = nn.ModuleList([nn.Conv2d(1027, 512, 1), nn.Conv2d(512, 512, 1), nn.Conv2d(512, 512, 1)])
        self.weightnet1, self.weightnet2 = _WeightNet(), _WeightNet()


class _Head(nn.Module):           # FlowHead / MotionHead (radarflow_util.py:240-285)
    def __init__(self, cout):
        super().__init__()
        self.sf_mlp = nn.ModuleList()
        last = 512
        for co in (256, 128, 64):
            self.sf_mlp.append(nn.Sequential(nn.Conv2d(last, co, 1, bias=False), nn.BatchNorm2d(co), nn.ReLU(inplace=False)))
            last = co
        self.conv2 = nn.Conv2d(64, cout, 1, bias=False)


class _FlowDecoder(nn.Module):    # FlowDecoder (radarflow_util.py:321-337): parameter container of RaFlow.fd_layer
    def __init__(self):
        super().__init__()
        self.mse = _MultiScale(1027, (512, 256, 64), (64, 64, 64))
        self.fp = _Head(3)


def _check_inputs(pc1, pc2, feature1, feature2, host=False):
    """The engine reads raw pointers: pc1 / feature1 must be (B,3,N1) and pc2 / feature2 (B,3,N2) with one B (N1 != N2 is fine: the
    reference's evaluation loop feeds un-resampled clouds, dataset/vod.py:92-93), fp32, contiguous, all on the same side (device or host).
    Returns the four tensors ready to pass."""
    ins = []
    for name, t in (("pc1", pc1), ("pc2", pc2), ("feature1", feature1), ("feature2", feature2)):
        if not torch.is_tensor(t):
            raise CmfError(f"{name} must be a tensor")
        if t.is_cuda == host:
            raise CmfError(f"{name}: " + ("forward_host takes host tensors" if host else "cmflow_b200 has no CPU path: inputs must be CUDA tensors"))
        ins.append(t.float().contiguous())
    s1, s2 = tuple(ins[0].shape), tuple(ins[1].shape)
    if len(s1) != 3 or s1[1] != 3:
        raise CmfError(f"pc1 must be (B,3,N), got {s1}")
    if len(s2) != 3 or s2[1] != 3 or s2[0] != s1[0]:
        raise CmfError(f"pc2 must be ({s1[0]},3,N2), got {s2}")
    if tuple(ins[2].shape) != s1:
        raise CmfError(f"feature1 has shape {tuple(ins[2].shape)}, expected {s1} like pc1")
    if tuple(ins[3].shape) != s2:
        raise CmfError(f"feature2 has shape {tuple(ins[3].shape)}, expected {s2} like pc2")
    if not host:
        for name, t in zip(("pc2", "feature1", "feature2"), ins[1:]):
            if t.device != ins[0].device:
                raise CmfError(f"{name} is on {t.device}, pc1 on {ins[0].device}")
    return ins


def _check_gfeat(gfeat, B, host, device=None):
    if gfeat is None:
        return None
    g = gfeat.float().contiguous()
    if tuple(g.shape) != (B, 256):
        raise CmfError(f"gfeat must be ({B},256), got {tuple(g.shape)}")
    if g.is_cuda == host or (not host and g.device != device):
        raise CmfError("gfeat must live where the inputs live")
    return g


class _EngineModel(nn.Module):
    _temporal = False
    _raflow = False

    def __init__(self, args):
        super().__init__()
        self.npoints = args.num_points
        self.mse_layer = _MultiScale(3, (32, 32, 64), (64, 64, 64))
        self.fc_layer = _Correlator()
        if self._raflow:
            self.rigid_thres, self.rigid_pcs = args.rigid_thres, 0.25           # raflow.py:16-17
            self.stat_thres = 0.5                                               # unused
            self.fd_layer = _FlowDecoder()
        else:
            self.stat_thres = 0.5 if self._temporal else args.stat_thres        # cmflow_t.py:18 hard-codes 0.50
            self.mse_layer2 = _MultiScale(1027, (512, 256, 64), (64, 64, 64))
            if self._temporal:
                self.gru = nn.GRU(input_size=256, hidden_size=256, num_layers=1)
            self.fp, self.mp = _Head(3), _Head(1)
        self._handle = None
        self.eval()

    # ---- engine lifecycle ------------------------------------------------------------------------
    def _drop_engine(self):
        if getattr(self, "_handle", None):
            lib().cmf_model_destroy(self._handle)
        self._handle = None

    def load_state_dict(self, *a, **k):
        self._drop_engine()
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._drop_engine()
        return super()._apply(fn, *a, **k)

    def __del__(self):
        try:
            self._drop_engine()
        except Exception:
            pass

    def _engine(self, device):
        if self._handle is None:
            blob = _weights.pack(self.state_dict(), temporal=self._temporal, raflow=self._raflow)
            L = lib()
            assert blob.size == L.cmf_model_blob_floats(int(self._temporal))
            h = ctypes.c_void_p()
            with torch.cuda.device(device):
                check(L.cmf_model_create(ctypes.byref(h), blob.ctypes.data_as(ctypes.c_void_p), blob.size,
                                         int(self._temporal), float(self.stat_thres)))
            self._handle = h
            if self._raflow:
                check(L.cmf_model_set_raflow(h, float(self.rigid_thres), float(self.rigid_pcs)))
            if getattr(self, "_mode", None) is not None:
                check(L.cmf_model_set_mode(h, self._mode))
        return self._handle

    def launches_per_forward(self):
        return lib().cmf_model_launches_per_forward(self._handle) if self._handle else 0

    def set_precision(self, mode):
        """'fp32' = strict fp32 FMA kernels (parity build); 'tf32x3' / 'fp16x3' = tcgen05 tensor cores with a 22-bit hi/lo operand
        split (3 MMAs per product, fp32 accumulate): kind::tf32, or kind::f16 at twice the rate with power-of-two operand scaling."""
        self._mode = {"fp32": 0, "tf32x3": 1, "fp16x3": 2}[mode]
        if self._handle is not None:
            check(lib().cmf_model_set_mode(self._handle, self._mode))

    def set_profiling(self, enable):
        check(lib().cmf_model_set_profiling(self._handle, int(bool(enable))))

    def read_profile(self):
        """{category: (device ms, launches, algorithmic FLOPs)} of the last forward (needs set_profiling(True))."""
        L = lib()
        n = L.cmf_model_profile_categories()
        ms = (ctypes.c_float * n)(); cnt = (ctypes.c_int * n)(); work = (ctypes.c_double * n)()
        check(L.cmf_model_read_profile(self._handle, ms, cnt, work))
        return {L.cmf_model_profile_name(i).decode(): (ms[i], cnt[i], work[i]) for i in range(n)}

    def workspace_bytes(self):
        return lib().cmf_model_workspace_bytes(self._handle) if self._handle else 0

    def tap(self, name, shape, dtype=torch.float32):
        """Copy of an engine workspace buffer of the last forward chunk (tests only)."""
        p = lib().cmf_model_tap(self._handle, name.encode())
        if not p:
            raise KeyError(name)

        class _Raw:          # zero-copy view of a raw device pointer through the CUDA array interface
            __cuda_array_interface__ = {"shape": tuple(shape), "typestr": {torch.float32: "<f4", torch.int32: "<i4"}[dtype],
                                        "data": (int(p), False), "version": 2}

        torch.cuda.synchronize()
        return torch.as_tensor(_Raw(), device="cuda").clone()

    def _warn_training(self):
        if self.training and not getattr(self, "_warned_training", False):
            self._warned_training = True
            warnings.warn("cmflow_b200 engines run inference only: BatchNorm uses its running statistics and no gradients flow, "
                          "even after .train()", RuntimeWarning, stacklevel=3)

    def _run(self, pc1, pc2, feature1, feature2, label_m, mode, gfeat):
        self._warn_training()
        ins = _check_inputs(pc1, pc2, feature1, feature2)
        B, _, N = ins[0].shape
        dev = ins[0].device
        h = self._engine(dev)
        sf = torch.empty(B, 3, N, device=dev); cls = torch.empty(B, 1, N, device=dev)
        T = torch.empty(B, 4, 4, device=dev); mask = torch.empty(B, N, dtype=torch.uint8, device=dev)
        gout = torch.empty(B, 256, device=dev) if self._temporal else None
        gprev = _check_gfeat(gfeat, B, False, dev) if self._temporal else None
        lab = None
        if mode == 'train' and label_m is not None:
            # models/cmflow.py:181-182 (cmflow_t.py:196-197): the pseudo motion labels replace the predicted scores in the ego-motion
            # head and in the refinement mask; stat_cls (returned) stays the network's own
            lab = label_m.to(dev).float().contiguous()
            if lab.numel() != B * N:
                raise CmfError(f"label_m must hold B*N = {B * N} values, got {tuple(label_m.shape)}")
        with torch.cuda.device(dev):
            check(lib().cmf_model_forward2(h, B, N, ins[1].shape[2], dptr(ins[0]), dptr(ins[1]), dptr(ins[2]), dptr(ins[3]), dptr(gprev), dptr(lab),
                                           dptr(sf), dptr(cls), dptr(T), dptr(mask), dptr(gout), stream_ptr()))
        return sf, cls, T, mask.bool(), gout

    def _host_args(self, pc1, pc2, feature1, feature2, gfeat, out):
        self._warn_training()
        ins = _check_inputs(pc1, pc2, feature1, feature2, host=True)
        B, _, N = ins[0].shape
        dev = torch.device("cuda", torch.cuda.current_device())
        h = self._engine(dev)
        want = {"sf_agg": ((B, 3, N), torch.float32), "stat_cls": ((B, 1, N), torch.float32), "pre_trans": ((B, 4, 4), torch.float32),
                "mask": ((B, N), torch.uint8), "gfeat": ((B, 256), torch.float32)}
        if out is None:
            out = {k: torch.empty(shp, dtype=dt).pin_memory() for k, (shp, dt) in want.items()}
        else:
            for k, (shp, dt) in want.items():
                t = out.get(k)
                if t is None or t.is_cuda or tuple(t.shape) != shp or t.dtype != dt or not t.is_contiguous():
                    raise CmfError(f"out[{k!r}] must be a contiguous host tensor of shape {shp}, dtype {dt}")
        hp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
        gprev = _check_gfeat(gfeat, B, True) if self._temporal else None
        args = (B, N, ins[1].shape[2], hp(ins[0]), hp(ins[1]), hp(ins[2]), hp(ins[3]), hp(gprev),
                hp(out["sf_agg"]), hp(out["stat_cls"]), hp(out["pre_trans"]), hp(out["mask"]), hp(out["gfeat"]), stream_ptr())
        return h, args, out, (ins, gprev)

    def forward_host(self, pc1, pc2, feature1, feature2, gfeat=None, out=None):
        """End-to-end call on HOST tensors (pinned for full speed): H2D, forward, D2H, synchronise
        (cmf_model_forward_host2).  Returns CPU tensors; pass `out` (dict of pinned tensors) to reuse buffers."""
        h, args, out, _keep = self._host_args(pc1, pc2, feature1, feature2, gfeat, out)
        check(lib().cmf_model_forward_host2(h, *args))
        return out

    def submit_host(self, slot, pc1, pc2, feature1, feature2, gfeat=None, out=None):
        """Pipelined form of forward_host (cmf_model_submit_host): enqueue upload, kernels and download of one call on staging slot 0 or 1
        and return at once; wait_host(slot) blocks until `out` is filled.  Alternating the two slots overlaps the copies of neighbouring
        calls with the kernels.  Inputs and `out` must be pinned and stay untouched until the wait."""
        h, args, out, keep = self._host_args(pc1, pc2, feature1, feature2, gfeat, out)
        check(lib().cmf_model_submit_host(h, int(slot), *args))
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[int(slot)] = keep                  # the host input tensors must outlive the asynchronous copies
        return out

    def wait_host(self, slot):
        check(lib().cmf_model_wait_host(self._handle, int(slot)))
        getattr(self, "_inflight", {}).pop(int(slot), None)


class CMFlow(_EngineModel):
    """models/cmflow.py:9 -- forward(pc1, pc2, feature1, feature2, label_m, mode) -> (sf_agg, stat_cls, pre_trans, mask)."""
    _temporal = False

    def forward(self, pc1, pc2, feature1, feature2, label_m, mode):
        sf, cls, T, mask, _ = self._run(pc1, pc2, feature1, feature2, label_m, mode, None)
        return sf, cls, T, mask


class CMFlow_T(_EngineModel):
    """models/cmflow_t.py:10 -- forward(..., label_m, mode, gfeat) -> (sf_agg, stat_cls, pre_trans, mask, gfeat)."""
    _temporal = True

    def forward(self, pc1, pc2, feature1, feature2, label_m, mode, gfeat):
        return self._run(pc1, pc2, feature1, feature2, label_m, mode, gfeat)


class RaFlow(_EngineModel):
    """models/raflow.py:9 -- forward(pc1, pc2, feature1, feature2, interval) -> (output, sf_agg, pre_trans, mask_s)."""
    _raflow = True

    def forward(self, pc1, pc2, feature1, feature2, interval):
        self._warn_training()
        ins = _check_inputs(pc1, pc2, feature1, feature2)
        B, _, N = ins[0].shape
        dev = ins[0].device
        dt = interval.to(dev).float().contiguous().view(-1)
        if dt.numel() != B:
            raise CmfError("interval must hold one value per frame pair")
        h = self._engine(dev)
        out = torch.empty(B, 3, N, device=dev); sf = torch.empty(B, 3, N, device=dev)
        T = torch.empty(B, 4, 4, device=dev); mask = torch.empty(B, N, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().cmf_model_forward_raflow2(h, B, N, ins[1].shape[2], dptr(ins[0]), dptr(ins[1]), dptr(ins[2]), dptr(ins[3]), dptr(dt),
                                                 dptr(out), dptr(sf), dptr(T), dptr(mask), stream_ptr()))
        return out, sf, T, mask.bool()
