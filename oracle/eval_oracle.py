"""numpy restatement of the reference's evaluation metrics (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows utils/eval_util.py:4-113 and utils/odometry_util.py:34-159 line by line (same dtypes: float32 inputs, float64 where the
reference's resolution vector promotes).  Pinned by tests/golden/eval_metrics.pt, produced by the UNMODIFIED reference functions
(tests/golden/make_golden.py --eval-only)."""
import numpy as np

LIDAR = (0.04, 0.4 * np.pi / 180, 0.08 * np.pi / 180)            # eval_util.py:13-15


def cartesian_res(pc, res):
    """get_carterian_res, eval_util.py:4-40: pc (B,3,N) float32, res (r, theta, phi) -> (B,N,3) float64."""
    res = np.array(res)
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    theta = np.arcsin(z / r)
    phi = np.arctan2(y, x)
    gx = np.stack((np.cos(phi) * np.cos(theta), -r * np.sin(theta) * np.cos(phi), -r * np.cos(theta) * np.sin(phi)), axis=2)
    gy = np.stack((np.sin(phi) * np.cos(theta), -r * np.sin(phi) * np.sin(theta), r * np.cos(theta) * np.cos(phi)), axis=2)
    gz = np.stack((np.sin(theta), r * np.cos(theta), np.zeros((np.size(x, 0), np.size(x, 1)))), axis=2)
    return np.stack((np.sum(abs(gx) * res, axis=2), np.sum(abs(gy) * res, axis=2), np.sum(abs(gz) * res, axis=2)), axis=2)


def eval_scene_flow(pc, pred, labels, mask, radar_res):
    """eval_util.py:42-83.  numpy float32 arrays: pc (B,3,N), pred / labels (B,N,3), mask (B,N)."""
    error = np.sqrt(np.sum((pred - labels) ** 2, 2) + 1e-20)
    gtflow_len = np.sqrt(np.sum(labels * labels, 2) + 1e-20)
    n = np.size(pred, 0) * np.size(pred, 1)
    epe = np.mean(error)
    accs = np.sum(np.logical_or(error <= 0.05, error / gtflow_len <= 0.05)) / n
    accr = np.sum(np.logical_or(error <= 0.10, error / gtflow_len <= 0.10)) / n
    res_r = np.sqrt(np.sum(cartesian_res(pc, (radar_res["r_res"], radar_res["theta_res"], radar_res["phi_res"])), 2) + 1e-20)
    res_l = np.sqrt(np.sum(cartesian_res(pc, LIDAR), 2) + 1e-20)
    re_error = error / (res_r / res_l)
    rne = np.mean(re_error)
    mov_rne = np.sum(re_error[mask == 0]) / (np.sum(mask == 0) + 1e-6)
    stat_rne = np.mean(re_error[mask == 1])
    sas = np.sum(np.logical_or(re_error <= 0.10, re_error / gtflow_len <= 0.10)) / n
    ras = np.sum(np.logical_or(re_error <= 0.20, re_error / gtflow_len <= 0.20)) / n
    return {"rne": rne, "50-50 rne": (mov_rne + stat_rne) / 2, "mov_rne": mov_rne, "stat_rne": stat_rne, "sas": sas, "ras": ras,
            "epe": epe, "accs": accs, "accr": accr}


def eval_motion_seg(pre, gt):
    """eval_util.py:99-113."""
    tp = np.logical_and(pre == 1, gt == 1).sum(); tn = np.logical_and(pre == 0, gt == 0).sum()
    fp = np.logical_and(pre == 1, gt == 0).sum(); fn = np.logical_and(pre == 0, gt == 1).sum()
    return {"acc": (tp + tn) / (tp + tn + fp + fn), "miou": 0.5 * (tp / (tp + fp + fn + 1e-10) + tn / (tn + fp + fn + 1e-10)),
            "sen": tp / (tp + fn + 1e-10)}


def eval_trans_rpe(gt_trans, pred_trans):
    """eval_util.py:86-97 with odometry_util.py:62-117 (E = gt^-1 pred), :133-138 (|t(E)|, rotation angle of E in degrees).
    The angle is |rotvec| of the rotation nearest to E's 3x3 block (scipy's from_matrix orthonormalises by SVD)."""
    rte, rae = [], []
    for G, Pm in zip(gt_trans.astype(np.float64), pred_trans.astype(np.float64)):
        Rinv = G[:3, :3].T
        E = np.eye(4)
        E[:3, :3] = Rinv @ Pm[:3, :3]
        E[:3, 3] = Rinv @ Pm[:3, 3] - Rinv @ G[:3, 3]
        U, _, Vt = np.linalg.svd(E[:3, :3])
        Rn = U @ Vt
        v = 0.5 * np.array([Rn[2, 1] - Rn[1, 2], Rn[0, 2] - Rn[2, 0], Rn[1, 0] - Rn[0, 1]])
        rae.append(np.degrees(np.arctan2(np.linalg.norm(v), 0.5 * (np.trace(Rn) - 1.0))))
        rte.append(np.linalg.norm(E[:3, 3]))
    return {"RTE": float(np.mean(rte)), "RAE": float(np.mean(rae))}
