"""Evaluation metrics on the device: host-side mirror of what the reference's
`Trainer.evaluate` (train.py:336-482) does with the persons returned by `Model.forward`, over the C-ABI entry points
`mhmr_eval_match_2d` / `mhmr_eval_points_error` (csrc/metrics.cu).  The matched pairs stay on the device between the
two kernels; the host reads back three small tensors per image."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import c_int, c_void_p, check, ptr


class AverageMeter:
    """utils/training.py:196-221."""

    def __init__(self, name):
        self.name, self.sum, self.count, self.val, self.avg = name, 0.0, 0, 0.0, 0.0

    def update(self, val, n=1):
        self.val = float(val)
        self.sum += float(val) * n
        self.count += n
        self.avg = self.sum / self.count


def compute_prf1(count, miss, fp):
    """utils/training.py:9-23 (host arithmetic on three integers)."""
    if count == 0:
        return 0, 0, 0
    tp, fn = count - miss, miss
    if tp == 0:
        return 0.0, 0.0, 0.0
    f1 = round(tp / (tp + 0.5 * (fp + fn)), 2)
    recall = round(tp / (tp + fn), 2)
    precision = round(tp / (tp + fp), 2)
    return 100.0 * precision, 100.0 * recall, 100.0 * f1


def _stream(dev):
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def match_2d_greedy(pred_j2d: torch.Tensor, gt_j2d: torch.Tensor, valid_mask: torch.Tensor | None = None,
                    iou_thresh: float = 0.05):
    """utils/training.py:25-147 on the device.  pred_j2d [P,J,2], gt_j2d [G,J,2] (CUDA, fp32).  Returns device tensors
    (pairs [min(P,G),2] int32, n_pairs [1] int32, pred_to_gt [P] int32, gt_to_pred [G] int32)."""
    lib = _lib.load()
    dev = gt_j2d.device
    assert dev.type == "cuda", "the evaluation kernels need CUDA tensors (no CPU fallback)"
    P, G, J = int(pred_j2d.shape[0]), int(gt_j2d.shape[0]), int(gt_j2d.shape[1])
    pred = pred_j2d.to(dev, torch.float32)[:, :J].contiguous() if P else torch.zeros(0, J, 2, device=dev)
    gt = gt_j2d.to(torch.float32).contiguous()
    vm = valid_mask.to(dev, torch.uint8).contiguous() if valid_mask is not None else None
    pairs = torch.full((max(1, min(P, G)), 2), -1, device=dev, dtype=torch.int32)
    n_pairs = torch.zeros(1, device=dev, dtype=torch.int32)
    p2g = torch.full((max(P, 1),), -1, device=dev, dtype=torch.int32)
    g2p = torch.full((G,), -1, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        check(lib.mhmr_eval_match_2d(ptr(pred), ptr(gt), ptr(vm), c_int(P), c_int(G), c_int(J),
                                     ctypes.c_float(iou_thresh), ptr(pairs), ptr(n_pairs), ptr(p2g), ptr(g2p),
                                     _stream(dev)), "mhmr_eval_match_2d")
    return pairs, n_pairs, p2g[:P], g2p


def points_error(pred, gt, pairs, n_pairs, pred_center=None, gt_center=None):
    """Mean point error (mm) and Procrustes-aligned mean point error (mm) of the matched pairs
    (train.py:387-394, :419-427).  pred [P,n,3], gt [G,n,3]; returns two device tensors [pairs.shape[0]]."""
    lib = _lib.load()
    dev = gt.device
    n = int(gt.shape[1])
    assert pred.shape[1] == n, "prediction and ground truth need the same number of points"
    M = int(pairs.shape[0])
    c = lambda t: None if t is None else t.to(dev, torch.float32).reshape(-1, 3).contiguous()
    pred, gt = pred.to(dev, torch.float32).contiguous(), gt.to(torch.float32).contiguous()
    pc, gc = c(pred_center), c(gt_center)
    err = torch.zeros(M, device=dev)
    pa = torch.zeros(M, device=dev)
    with torch.cuda.device(dev):
        check(lib.mhmr_eval_points_error(ptr(pred), ptr(pc), ptr(gt), ptr(gc), ptr(pairs), ptr(n_pairs), c_int(M),
                                         c_int(n), ptr(err), ptr(pa), _stream(dev)), "mhmr_eval_points_error")
    return err, pa


class Evaluator:
    """The accumulation loop of `Trainer.evaluate` (train.py:336-482) for predictions in the reference's person-dict
    format (model.py:329-347) and ground truths {j2d [G,J,2], v3d [G,V,3], transl_pelvis [G,1,3]}."""

    def __init__(self):
        self.meters = {k: AverageMeter(k) for k in ("pve", "pa_pve", "precision", "recall", "f1_score")}
        self.count = self.miss = self.fp = 0

    def update(self, persons: list, gt: dict):
        G = int(gt["j2d"].shape[0])
        dev = gt["j2d"].device
        if len(persons):
            pj = torch.stack([p["j2d"] for p in persons])
            pv = torch.stack([p["v3d"] for p in persons])
            pp = torch.stack([p["transl_pelvis"].reshape(3) for p in persons])
        else:
            pj, pv, pp = torch.zeros(0, gt["j2d"].shape[1], 2, device=dev), None, None
        pairs, n_pairs, p2g, g2p = match_2d_greedy(pj, gt["j2d"])
        if len(persons):
            pve, pa = points_error(pv, gt["v3d"], pairs, n_pairs, pp, gt["transl_pelvis"])
        n = int(n_pairs.item())  # the one host read-back of the image
        self.count += G
        self.miss += G - n
        self.fp += len(persons) - n
        if n:
            for a, b in zip(pve[:n].tolist(), pa[:n].tolist()):
                self.meters["pve"].update(a)
                self.meters["pa_pve"].update(b)
        return pairs[:n]

    def summary(self) -> dict:
        precision, recall, f1 = compute_prf1(self.count, self.miss, self.fp)
        out = {k: m.avg for k, m in self.meters.items() if k in ("pve", "pa_pve")}
        out.update(precision=precision, recall=recall, f1_score=f1)
        return out
