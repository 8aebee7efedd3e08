"""The reference's evaluation metrics (utils/eval_util.py:42-117; eval loop main_util.py:106-203) on the GPU.

`eval_scene_flow`, `eval_motion_seg` and `eval_trans_RPE` keep the reference's names, arguments and returned dictionaries, but
the work is one reduction kernel per call on the tensors where they already live (cmf_eval_*_sums, csrc/eval_kernels.cu) and one
read-back of a dozen doubles -- the reference copies pc / pred / labels / mask to the host and runs numpy (+ a scipy rotation-vector
conversion per pair) for every batch.  `EvalAccumulator` is the loop body of main_util.py:175-203 without ANY per-batch host
synchronisation: per-batch metrics are formed on the device from the sums, accumulated weighted by batch size as the reference does,
and read back once at the end -- after one all-reduce when the evaluation is sharded over ranks (the only collective of the multi-GPU
path).  There is no CPU path.
"""
import torch

from ._lib import CmfError, check, dptr, lib, stream_ptr

SF_KEYS = ("rne", "50-50 rne", "mov_rne", "stat_rne", "sas", "ras", "epe", "accs", "accr")      # main_util.py:108-109
SEG_KEYS = ("acc", "miou", "sen")
POSE_KEYS = ("RTE", "RAE")
LIDAR_RES = (0.04, 0.4 * 3.141592653589793 / 180, 0.08 * 3.141592653589793 / 180)               # eval_util.py:13-15


def _f32(t, name):
    if not t.is_cuda:
        raise CmfError(f"cmflow_b200.eval_util has no CPU path: {name} must be a CUDA tensor")
    return t.detach().float().contiguous()


def scene_flow_metrics_device(pc, pred, labels, mask, radar_res):
    """(9,) float64 device tensor in SF_KEYS order for one batch: pc (B,3,N), pred / labels (B,N,3), mask (B,N) (1 = static)."""
    pc, pred, labels, mask = _f32(pc, "pc"), _f32(pred, "pred"), _f32(labels, "labels"), _f32(mask, "mask")
    B, _, N = pc.shape
    s = torch.zeros(11, dtype=torch.float64, device=pc.device)
    with torch.cuda.device(pc.device):
        check(lib().cmf_eval_scene_flow_sums(B, N, dptr(pc), dptr(pred), dptr(labels), dptr(mask), float(radar_res["r_res"]),
                                             float(radar_res["theta_res"]), float(radar_res["phi_res"]), dptr(s), stream_ptr()))
    cnt = s[0]
    rne = s[4] / cnt                                  # eval_util.py:69
    mov = s[5] / (s[6] + 1e-6)                        # :70
    stat = s[7] / s[8]                                # :71 (np.mean of an empty selection is NaN there too)
    return torch.stack([rne, (mov + stat) / 2, mov, stat, s[9] / cnt, s[10] / cnt, s[1] / cnt, s[2] / cnt, s[3] / cnt])


def motion_seg_metrics_device(pre, gt):
    pre, gt = _f32(pre, "pre"), _f32(gt, "gt")
    c = torch.zeros(4, dtype=torch.float64, device=pre.device)
    with torch.cuda.device(pre.device):
        check(lib().cmf_eval_motion_seg_counts(pre.numel(), dptr(pre), dptr(gt), dptr(c), stream_ptr()))
    tp, tn, fp, fn = c[0], c[1], c[2], c[3]
    return torch.stack([(tp + tn) / (tp + tn + fp + fn), 0.5 * (tp / (tp + fp + fn + 1e-10) + tn / (tn + fp + fn + 1e-10)),
                        tp / (tp + fn + 1e-10)])      # eval_util.py:107-109


def pose_metrics_device(gt_trans, rigid_trans):
    gt_trans, rigid_trans = _f32(gt_trans, "gt_trans"), _f32(rigid_trans, "rigid_trans")
    s = torch.zeros(3, dtype=torch.float64, device=gt_trans.device)
    with torch.cuda.device(gt_trans.device):
        check(lib().cmf_eval_rpe_sums(gt_trans.shape[0], dptr(gt_trans), dptr(rigid_trans), dptr(s), stream_ptr()))
    return torch.stack([s[1] / s[0], s[2] / s[0]])    # eval_util.py:93-94


def eval_scene_flow(pc, pred, labels, mask, args):
    """utils/eval_util.py:42-83 -- same arguments (args.radar_res), same dictionary."""
    return dict(zip(SF_KEYS, scene_flow_metrics_device(pc, pred, labels, mask, args.radar_res).tolist()))


def eval_motion_seg(pre, gt):
    """utils/eval_util.py:99-113."""
    return dict(zip(SEG_KEYS, motion_seg_metrics_device(pre, gt).tolist()))


def eval_trans_RPE(gt_trans, rigid_trans):
    """utils/eval_util.py:86-97."""
    return dict(zip(POSE_KEYS, pose_metrics_device(gt_trans, rigid_trans).tolist()))


class EvalAccumulator:
    """main_util.py:106-203: metric sums weighted by batch size, divided by the number of evaluated pairs at the end."""

    def __init__(self, radar_res, device):
        self.radar_res = radar_res
        self.acc = torch.zeros(len(SF_KEYS) + len(SEG_KEYS) + len(POSE_KEYS) + 1, dtype=torch.float64, device=device)

    def add(self, pc1, pred_f, gt, mask, pred_m, gt_trans, pred_trans):
        """pc1 (B,3,N); pred_f (B,3,N) as the models return it (transposed here like main_util.py:175); gt (B,N,3); mask, pred_m (B,N)."""
        B = pc1.shape[0]
        m = torch.cat([scene_flow_metrics_device(pc1, pred_f.transpose(2, 1), gt, mask, self.radar_res),
                       motion_seg_metrics_device(pred_m, mask), pose_metrics_device(gt_trans, pred_trans),
                       torch.ones(1, dtype=torch.float64, device=self.acc.device)])
        self.acc += B * m                              # main_util.py:176-195

    def result(self, all_reduce=True):
        acc = self.acc.clone()
        if all_reduce and torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(acc)          # the only collective of a sharded evaluation
        v = (acc[:-1] / acc[-1]).tolist()              # main_util.py:197-202
        a, b = len(SF_KEYS), len(SF_KEYS) + len(SEG_KEYS)
        return dict(zip(SF_KEYS, v[:a])), dict(zip(SEG_KEYS, v[a:b])), dict(zip(POSE_KEYS, v[b:])), int(round(acc[-1].item()))
