"""CPU restatement of the evaluation metrics of the reference's `Trainer.evaluate` (train.py:336-482) and of its
matching helpers (utils/training.py:9-193): SURVEY.md §8(f) row 3.  numpy / torch, every function cites the
reference lines it follows.  `match_2d_greedy` / `get_bbx_overlap` / `compute_prf1` are pinned against the
reference's OWN functions by oracle/make_golden.py (tests/golden/eval_matching.npz); the Procrustes alignment
comes from roma_ref (third-party, second-sourced in tests/test_oracle_second_source.py).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import numpy as np
import torch

from . import roma_ref


def compute_prf1(count, miss, fp):
    """utils/training.py:9-23."""
    if count == 0:
        return 0, 0, 0
    tp, fn = count - miss, miss
    if tp == 0:
        return 0.0, 0.0, 0.0
    f1 = round(tp / (tp + 0.5 * (fp + fn)), 2)
    recall = round(tp / (tp + fn), 2)
    precision = round(tp / (tp + fp), 2)
    return 100.0 * precision, 100.0 * recall, 100.0 * f1


def get_bbx_overlap(p1, p2):
    """utils/training.py:149-193: IoU of the keypoint bounding boxes, areas with the +1 pixel convention."""
    mn1, mn2, mx1, mx2 = p1.min(0), p2.min(0), p1.max(0), p2.max(0)
    x_left, y_top = max(mn1[0], mn2[0]), max(mn1[1], mn2[1])
    x_right, y_bottom = min(mx1[0], mx2[0]), min(mx1[1], mx2[1])
    inter = max(0, x_right - x_left + 1) * max(0, y_bottom - y_top + 1)
    a1 = (mx1[0] - mn1[0] + 1) * (mx1[1] - mn1[1] + 1)
    a2 = (mx2[0] - mn2[0] + 1) * (mx2[1] - mn2[1] + 1)
    return inter / float(a1 + a2 - inter)


def pair_error(pred, gt, vmask):
    """utils/training.py:50: np.linalg.norm(D, 2) of the [J, 2] difference — ord=2 of a MATRIX is its largest
    singular value (not the Frobenius norm)."""
    d = pred[vmask, :2] - gt[vmask, :2]
    return np.linalg.norm(d, 2)


def match_2d_greedy(pred_kps, gtkp, valid_mask, iou_thresh=0.05):
    """utils/training.py:25-147 with valid=None (the only way train.py:364 calls it).  Returns
    (bestMatch [n, 2] (pred, gt) in the order they were found, falsePositives, misses)."""
    P, G = len(pred_kps), len(gtkp)
    err = np.full((P * G,), np.inf)
    for p in range(P):
        for g in range(G):
            err[p * G + g] = pair_error(pred_kps[p], gtkp[g], valid_mask[g])
    gt_assigned = np.zeros(G, dtype=bool)
    op_assigned = np.zeros(P, dtype=bool)
    best, fp_counter = [], 0
    while gt_assigned.sum() < G and op_assigned.sum() + fp_counter < P:
        found = false_positive = False
        p = g = -1
        while not found:
            if np.all(np.isinf(err)):
                raise RuntimeError("match_2d_greedy: no candidate pair left (the reference loops forever here)")
            i = int(np.argmin(err))
            p, g = divmod(i, G)
            iou = get_bbx_overlap(pred_kps[p], gtkp[g])
            err[i] = np.inf
            if not op_assigned[p] and not gt_assigned[g] and iou >= iou_thresh:
                found = True
            elif iou < iou_thresh:
                found = false_positive = True
                fp_counter += 1
        if not false_positive:
            best.append((p, g))
            op_assigned[p] = gt_assigned[g] = True
    best = np.array(best, dtype=np.int64).reshape(-1, 2)
    false_positives = [p for p in range(P) if p not in set(best[:, 0].tolist())]
    misses = [g for g in range(G) if g not in set(best[:, 1].tolist())]
    return best, false_positives, misses


def points_errors(pred, gt):
    """train.py:387-394 (PVE, PA-PVE) and :419-427 (MPJPE, PA-MPJPE) for already centred point sets [n, 3]:
    mean Euclidean distance in mm, before and after the Procrustes (similarity) alignment of pred onto gt."""
    err = (torch.sqrt(((gt - pred) ** 2).sum(-1)) * 1000).mean()
    R, t, s = roma_ref.rigid_points_registration(pred, gt, compute_scaling=True)
    pa = s * (R.reshape(1, 3, 3) @ pred.reshape(-1, 3, 1)).reshape(-1, 3) + t
    pa_err = (torch.sqrt(((gt - pa) ** 2).sum(-1)) * 1000).mean()
    return err, pa_err


def evaluate_image(pred_j2d, pred_v3d, pred_pelvis, gt_j2d, gt_v3d, gt_pelvis):
    """One iteration of the loop of train.py:346-395 for a single image: matching on 2-D joints, then PVE /
    PA-PVE of the pelvis-centred meshes of the matched pairs.  Returns dict(best, fp, miss, pve[n], pa_pve[n])."""
    G = gt_j2d.shape[0]
    kp_gt = gt_j2d.numpy()
    kp_pred = np.asarray([p.numpy()[: kp_gt.shape[1]] for p in pred_j2d]) if len(pred_j2d) else np.zeros((0,) + kp_gt.shape[1:])
    best, fps, misses = match_2d_greedy(kp_pred, kp_gt, np.ones_like(kp_gt[..., 0]).astype(np.bool_))
    pve, pa = [], []
    for pid, gid in best:
        v = gt_v3d[gid] - gt_pelvis[gid].reshape(1, 3)
        vh = pred_v3d[pid] - pred_pelvis[pid].reshape(1, 3)
        e, pe = points_errors(vh, v)
        pve.append(e.item())
        pa.append(pe.item())
    return dict(best=best, fp=fps, miss=misses, pve=np.array(pve), pa_pve=np.array(pa), count=G)
